#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (+ reference arm), timelines, sweep, ncu launch list, ncu --set full captures.
# Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 200 python tools/timeline_fwd.py > gpurun_out/timeline_fwd.txt 2>&1
timeout 200 python tools/timeline_bwd.py > gpurun_out/timeline_bwd.txt 2>&1
timeout 200 python tools/timeline_tail.py 2>&1 | grep -v "^W\|OMP_NUM" > gpurun_out/timeline_tail_n1.txt
timeout 300 python tools/sweep_c3.py > gpurun_out/sweep_c3.jsonl 2> gpurun_out/sweep_c3.err
timeout 120 python tools/encode_c5.py > gpurun_out/encode_c5_n1.json 2> gpurun_out/encode_c5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log | cut -c1-300; cut -c1-700 gpurun_out/bench.json; tail -2 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_ref.json
bash tools/ncu_full.sh vqb_bwd_h2_kernel bwd
bash tools/ncu_full.sh vqb_fwd_tc_kernel fwd
bash tools/ncu_full.sh bwd_tail_h2_kernel tail
# config-3 kernels: the 1xTF32 search at K=8192, D=256 and the per-code gather-sum of the large-table scatter
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vqb_fwd_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_search \
    python tools/prof_search.py 8192 256 > gpurun_out/ncu_full_search.log 2>&1
VQB_SWEEP_POINTS="8192x256" VQB_SWEEP_N=1048576 timeout 600 ncu --set full --clock-control none --import-source on -k regex:segsum_kernel -s 2 -c 1 -f -o gpurun_out/prof_segsum \
    python tools/sweep_c3.py > gpurun_out/ncu_full_segsum.log 2>&1
ls -la gpurun_out/*.ncu-rep
