#!/bin/bash
# 2-GPU visit: fused-exchange check, then the bench at N=2 with the fused exchange and with NCCL.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29521 tools/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "dist_check exit $?" >> gpurun_out/dist_check.log
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/dist_check.log | tail -15
timeout 280 $TR --master-port 29522 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "exit $?" >> gpurun_out/bench_n2.err
VQB_NCCL_ALLREDUCE=1 timeout 280 $TR --master-port 29523 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err
timeout 200 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-330 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err; cut -c1-330 gpurun_out/bench_n2_nccl.json; cut -c1-330 gpurun_out/bench.json
timeout 300 python -m pytest tests -m gpu -q -x --timeout 280 2>&1 | tail -5
