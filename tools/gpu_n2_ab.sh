#!/bin/bash
# Same-box A/B of the N = 2 step: one GPU alone, two replicas without any exchange, the fused exchange (deferred / in-step), NCCL.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
show() { python -c "
import json,sys; d=json.load(open('$1')); print('$2', 'n', d['n_gpus'], 'ms/step %.4f' % d['ms_per_step'], 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['clocks'])"; }
timeout 200 $TR --master-port 29521 tools/dist_check.py > gpurun_out/dist_check_n2.log 2>&1; echo "dist_check exit $?"
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/dist_check_n2.log | tail -3 | cut -c1-700
timeout 200 python bench.py --steps 300 --warmup 10 --no-sweep > gpurun_out/ab_n1.json 2> gpurun_out/ab_n1.err; show gpurun_out/ab_n1.json n1
VQB_BENCH_NO_EXCHANGE=1 timeout 200 $TR --master-port 29531 bench.py --gpus 2 --steps 300 --warmup 10 > gpurun_out/ab_n2_noex.json 2> gpurun_out/ab_n2_noex.err; show gpurun_out/ab_n2_noex.json n2_no_exchange
timeout 200 $TR --master-port 29532 bench.py --gpus 2 --steps 300 --warmup 10 > gpurun_out/ab_n2_defer.json 2> gpurun_out/ab_n2_defer.err; show gpurun_out/ab_n2_defer.json n2_deferred
VQB_NO_DEFER=1 timeout 200 $TR --master-port 29533 bench.py --gpus 2 --steps 300 --warmup 10 > gpurun_out/ab_n2_instep.json 2> gpurun_out/ab_n2_instep.err; show gpurun_out/ab_n2_instep.json n2_in_step
[ -n "$WITH_NCCL" ] && { VQB_NCCL_ALLREDUCE=1 timeout 200 $TR --master-port 29534 bench.py --gpus 2 --steps 300 --warmup 10 > gpurun_out/ab_n2_nccl.json 2> gpurun_out/ab_n2_nccl.err; show gpurun_out/ab_n2_nccl.json n2_nccl; }
exit 0
