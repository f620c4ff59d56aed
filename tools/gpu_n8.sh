#!/bin/bash
# 8-GPU visit (one box): exchange check, the bench line at N = 1 / 8 on the SAME box (fused exchange, no exchange, NCCL) and
# the encode path (configs[4]) at N = 1 / 8.  Everything lands in gpurun_out/ (copied to profiles/r2_*_n8* by hand).
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
show() { python -c "
import json,sys; d=json.load(open('$1')); print('$2', 'n', d['n_gpus'], 'ms/step %.4f' % d['ms_per_step'], 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['clocks'])"; }
timeout 200 $TR --master-port 29521 tools/dist_check.py > gpurun_out/dist_check_n$N.log 2>&1; echo "dist_check exit $?"
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/dist_check_n$N.log | tail -2 | cut -c1-700
timeout 200 python bench.py --steps 300 --warmup 10 --no-sweep > gpurun_out/n8box_n1.json 2> gpurun_out/n8box_n1.err; show gpurun_out/n8box_n1.json n1_same_box
timeout 200 $TR --master-port 29522 bench.py --gpus $N --steps 300 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; show gpurun_out/bench_n$N.json fused_deferred
VQB_BENCH_NO_EXCHANGE=1 timeout 200 $TR --master-port 29523 bench.py --gpus $N --steps 300 --warmup 10 > gpurun_out/bench_n${N}_noex.json 2> gpurun_out/bench_n${N}_noex.err; show gpurun_out/bench_n${N}_noex.json no_exchange
VQB_NCCL_ALLREDUCE=1 timeout 200 $TR --master-port 29524 bench.py --gpus $N --steps 300 --warmup 10 > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err; show gpurun_out/bench_n${N}_nccl.json nccl
VQB_NO_DEFER=1 timeout 200 $TR --master-port 29526 bench.py --gpus $N --steps 300 --warmup 10 > gpurun_out/bench_n${N}_instep.json 2> gpurun_out/bench_n${N}_instep.err; show gpurun_out/bench_n${N}_instep.json fused_in_step
timeout 200 python bench.py --workload encode > gpurun_out/encode_n8box_n1.json 2> gpurun_out/encode_n8box_n1.err; cut -c1-160 gpurun_out/encode_n8box_n1.json
timeout 200 $TR --master-port 29525 bench.py --workload encode > gpurun_out/encode_n$N.json 2> gpurun_out/encode_n$N.err; echo "encode exit $?"; python -c "
import json; d=json.load(open('gpurun_out/encode_n$N.json')); print('encode n', d['n_gpus'], 'value %.4g' % d['value'], d['parity_mode'], d['clocks'])"
exit 0
