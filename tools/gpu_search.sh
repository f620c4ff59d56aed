#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 280 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | head -20
timeout 300 python tools/sweep_c3.py > gpurun_out/sweep_c3.jsonl 2> gpurun_out/sweep_c3.err
python -c "
import sys,json
for l in open('gpurun_out/sweep_c3.jsonl'):
    d=json.loads(l); print(d['K'],d['D'],'fwd %.3f ms %.1f TF (%.0f%% tf32) %.1f%% hbm rerank %d full %d | scatter %.3f ms %.1f%% hbm'%(d['fwd_ms'],d['search_tflops'],100*d['tensor_frac_of_tf32_peak'],100*d['fwd_hbm_frac'],d['reranked_rows'],d['full_scan_rows'],d['scatter_ms'],100*d['scatter_hbm_frac']))"
tail -3 gpurun_out/sweep_c3.err
