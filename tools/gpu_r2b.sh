#!/bin/bash
# Round 2 evidence visit (1 GPU): parity (whole GPU suite), timelines, both bench arms, the f3 / f4 A/Bs, launch list,
# ncu --set full of the two parity-mode kernels.  Everything lands in gpurun_out/; tools/collect_profiles.sh copies it.
mkdir -p gpurun_out; rm -f gpurun_out/test_records.jsonl
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|Error" gpurun_out/pytest_gpu.log | cut -c1-300 | head -20
timeout 200 python tools/timeline_pc.py > gpurun_out/timeline_pc.txt 2>&1; tail -4 gpurun_out/timeline_pc.txt | cut -c1-250
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cut -c1-400 gpurun_out/bench_ref.json
timeout 400 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print("ms/step %.4f" % d["ms_per_step"], d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"], d["roofline"].get("frac_per_kernel"), "e2e %.3g" % d["e2e"]["value"], d.get("clocks"))
    print("cpu_baseline", d.get("cpu_baseline"))
    for p in d.get("sweep", {}).get("points", []):
        print(p)
except Exception as e:
    print("unreadable:", e)
PY
timeout 200 python tools/ctc_fold_ab.py > gpurun_out/ctc_fold_ab.json 2>/dev/null; cat gpurun_out/ctc_fold_ab.json
timeout 200 python tools/length_aware_ab.py > gpurun_out/length_aware_ab.json 2>/dev/null; cat gpurun_out/length_aware_ab.json
timeout 200 python bench.py --workload encode > gpurun_out/encode_n1.json 2>/dev/null; cut -c1-200 gpurun_out/encode_n1.json
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-sweep > gpurun_out/ncu_bench.log 2>&1
bash tools/ncu_full.sh vqb_bwd_pcode_kernel bwd
bash tools/ncu_full.sh vqb_fwd_pcode_kernel fwd
fi
