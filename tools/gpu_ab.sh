#!/bin/bash
# A/B of the launch-chain options on one GPU + the full GPU test suite (no -x).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 280 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|Error|assert " gpurun_out/pytest_gpu.log | head -40
for v in "" "VQB_NO_PDL=1" "VQB_NO_TAIL=1" "VQB_NO_PDL=1 VQB_NO_TAIL=1"; do
  echo "== $v"
  env $v timeout 200 python bench.py --steps 200 --warmup 10 2> gpurun_out/ab.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['roofline']['kernel_ms'])"
done
