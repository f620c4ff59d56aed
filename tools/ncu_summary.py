#!/usr/bin/env python
"""Summarise `ncu --set full` reports (read offline, no GPU needed):
    python tools/ncu_summary.py ROUND gpurun_out/prof_fwd.ncu-rep gpurun_out/prof_bwd.ncu-rep ...
writes profiles/<ROUND>_ncu_full_<name>.csv (selected metrics per captured launch) and merges the per-launch DRAM
traffic of each kernel into profiles/ncu_traffic.json (read by bench.py for roofline.traffic)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__cycles_active.avg",
        "sm__pipe_tensor_subpipe_mma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "gpc__cycles_elapsed.max", "sm__cycles_active.avg"]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return v * mult


def main():
    rnd, reps = sys.argv[1], sys.argv[2:]
    traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.isfile(traffic_path) else {}
    for rep in reps:
        name = re.sub(r"^prof_", "", os.path.splitext(os.path.basename(rep))[0])
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print("no data in", rep); continue
        hdr, units, data = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(hdr)}
        keep = [k for k in KEEP if k in col]
        dst = os.path.join(ROOT, "profiles", "%s_ncu_full_%s.csv" % (rnd, name))
        with open(dst, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["kernel", "grid", "block"] + ["%s [%s]" % (k, units[col[k]]) for k in keep])
            for r in data:
                w.writerow([r[col["Kernel Name"]][:80], r[col.get("Grid Size", 0)], r[col.get("Block Size", 0)]] + [r[col[k]] for k in keep])
        print("wrote", dst)
        per = {}
        for r in data:
            kn = re.sub(r"^void\s+", "", r[col["Kernel Name"]])
            kn = re.split(r"[<(]", kn.replace("vqb::", ""))[0]
            rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            t = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
            tu = units[col["gpu__time_duration.sum"]]
            t_us = t * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(tu, 1)
            per.setdefault(kn, []).append((rd + wr, rd, wr, t_us))
        for kn, v in per.items():
            n = len(v)
            if name not in ("fwd", "bwd", "tail"):
                kn = "%s@%s" % (kn, name)                   # captures of other workloads never shadow the bench kernels
            traffic[kn] = {"dram_bytes_per_launch": sum(x[0] for x in v) / n, "dram_read": sum(x[1] for x in v) / n,
                           "dram_write": sum(x[2] for x in v) / n, "ncu_time_us": sum(x[3] for x in v) / n, "launches": n,
                           "source": "%s_ncu_full_%s.csv (ncu --set full --clock-control none, bench.py workload)" % (rnd, name)}
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
