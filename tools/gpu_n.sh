#!/bin/bash
# N-GPU visit (N = $1): fused-exchange check, then the bench at N with the fused exchange and with NCCL.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29521 tools/dist_check.py > gpurun_out/dist_check_n$N.log 2>&1; echo "dist_check exit $?" >> gpurun_out/dist_check_n$N.log
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/dist_check_n$N.log | tail -6
timeout 280 $TR --master-port 29522 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "exit $?" >> gpurun_out/bench_n$N.err
VQB_NCCL_ALLREDUCE=1 timeout 280 $TR --master-port 29523 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err
for f in gpurun_out/bench_n$N.json gpurun_out/bench_n${N}_nccl.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', d['n_gpus'], 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['config']['parallelism'][:90])"; done
tail -2 gpurun_out/bench_n$N.err
