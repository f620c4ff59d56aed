#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 280 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | head -30
for i in 1 2 3; do python tools/dbg_skip.py | grep -c "bad rows \[\]"; done
timeout 200 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
timeout 300 python tools/sweep_c3.py > gpurun_out/sweep_c3.jsonl 2> gpurun_out/sweep_c3.err
python -c "
import sys,json
for l in open('gpurun_out/sweep_c3.jsonl'):
    d=json.loads(l); print(d['K'],d['D'],'fwd %.3f ms %.1f TF (%.0f%% tf32) %.1f%% hbm | scatter %.3f ms %.1f%% hbm'%(d['fwd_ms'],d['search_tflops'],100*d['tensor_frac_of_tf32_peak'],100*d['fwd_hbm_frac'],d['scatter_ms'],100*d['scatter_hbm_frac']))"
tail -3 gpurun_out/sweep_c3.err
VQB_SWEEP_POINTS="256x64,8192x64,1024x256,8192x256" timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_sweep.csv python tools/sweep_c3.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_sweep.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    k=(r[4].split('(')[0][-60:], r[8]); agg.setdefault(k,[]).append(float(r[-1])/1e3)
for k,v in agg.items(): print('%-62s grid %-14s n=%3d  avg %9.1f us  min %9.1f' % (k[0],k[1],len(v),sum(v)/len(v),min(v)))
PY
