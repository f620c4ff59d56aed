#!/bin/bash
for v in "VQB_X=1" "VQB_GATHER_LDG=1" "VQB_X=2" "VQB_GATHER_LDG=1"; do
  env $v timeout 200 python bench.py --steps 300 --warmup 10 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$v', d['ms_per_step'], d['roofline']['kernel_ms'])"
done
VQB_GATHER_LDG=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "l2_module or config2" --timeout 200 2>&1 | tail -2
