#!/bin/bash
# Round 2 GPU visit: parity (whole GPU suite), timelines, bench, launch list, ncu --set full of the two parity-mode kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|Error" gpurun_out/pytest_gpu.log | cut -c1-300 | head -20
timeout 200 python tools/timeline_pc.py > gpurun_out/timeline_pc.txt 2>&1; cat gpurun_out/timeline_pc.txt | head -70
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
for f in ("bench",):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "ms/step %.4f" % d["ms_per_step"], d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"], "e2e %.3g" % d["e2e"]["value"], d.get("clocks"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
bash tools/ncu_full.sh vqb_bwd_pcode_kernel bwd
bash tools/ncu_full.sh vqb_fwd_pcode_kernel fwd
fi
