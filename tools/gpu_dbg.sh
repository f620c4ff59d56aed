#!/bin/bash
mkdir -p gpurun_out
for v in "VQB_X=1" "VQB_NO_PDL=1" "VQB_NO_TAIL_TEST=1" "VQB_NO_PDL=1 VQB_NO_TAIL_TEST=1"; do
  echo "== $v"
  env $v timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "skip_train" --timeout 100 2>&1 | grep -E "^E  |passed|failed" | head -8
done
for v in "VQB_X=1" "VQB_NO_PDL=1"; do
  echo "== bench $v"
  env $v timeout 200 python bench.py --steps 200 --warmup 10 2> gpurun_out/ab.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['roofline']['kernel_ms'])"
done
timeout 300 python -m pytest tests/test_gpu_segment.py -m gpu -q --timeout 100 2>&1 | tail -15
