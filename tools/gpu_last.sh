#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 280 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 200 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
VQB_SWEEP_POINTS="8192x64,4096x256,8192x256" timeout 300 python tools/sweep_c3.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['K'],d['D'],'fwd %.3f ms %.1f TF (%.0f%% tf32)'%(d['fwd_ms'],d['search_tflops'],100*d['tensor_frac_of_tf32_peak']))"
