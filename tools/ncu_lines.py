"""Per-source-line summary of an `ncu --set full --import-source on` report (offline): stall samples and executed warp
instructions per CUDA source line, heaviest first.   python tools/ncu_lines.py gpurun_out/prof_bwd.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, hdr, recs, seen_kernel = None, None, [], 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        seen_kernel += 1
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0] not in ("", "File Path") and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        try:
            recs.append((cur, int(r[0]), r[1].strip()[:110], int(d.get("# Samples", 0) or 0), int(d.get("Instructions Executed", 0) or 0), d))
        except ValueError:
            pass
# the report may hold several launches of the same kernel: keep the aggregate as printed (first function block only)
tot_s = sum(x[3] for x in recs) or 1
tot_i = sum(x[4] for x in recs) or 1
print("total samples %d, total warp instructions %d" % (tot_s, tot_i))
print("---- by stall samples")
for f, ln, src, s, i, d in sorted(recs, key=lambda x: -x[3])[:top]:
    st = sorted(((k, int(v)) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0), key=lambda kv: -kv[1])[:3]
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %s   %s" % (100.0 * s / tot_s, 100.0 * i / tot_i, f, ln, src, st))
print("---- by instructions")
for f, ln, src, s, i, d in sorted(recs, key=lambda x: -x[4])[:top]:
    print("%5.1f%% ins %5.1f%% smp  %s:%d  %s" % (100.0 * i / tot_i, 100.0 * s / tot_s, f, ln, src))
