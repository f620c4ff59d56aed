#!/bin/bash
# GPU visit for SURVEY 8f rank 3: the new test first (short timeout), the whole GPU suite, the A/B, one bench line.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "ctc_log_probs_from" > gpurun_out/pytest_f3.log 2>&1; echo "f3 exit $?"
tail -30 gpurun_out/pytest_f3.log | cut -c1-250
timeout 200 python tools/ctc_fold_ab.py > gpurun_out/ctc_fold_ab.json 2> gpurun_out/ctc_fold_ab.err; cat gpurun_out/ctc_fold_ab.json; tail -3 gpurun_out/ctc_fold_ab.err | cut -c1-300
VQB_GLOGP_NO_TMA=1 timeout 200 python tools/ctc_fold_ab.py 2>/dev/null | sed 's/^/no-tma: /'
[ -n "$QUICK" ] && exit 0
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|Error" gpurun_out/pytest_gpu.log | cut -c1-300 | head -20
timeout 300 python bench.py --steps 200 --warmup 10 --no-sweep > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print("ms/step %.4f" % d["ms_per_step"], d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"], "e2e %.3g" % d["e2e"]["value"], d.get("clocks"))
except Exception as e:
    print("unreadable:", e)
PY
