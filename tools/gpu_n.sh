#!/bin/bash
# N-GPU visit (N = $1): GPU test suite (the >= 2-GPU tests included), fused-exchange check, bench at N (fused exchange and NCCL),
# config 4 (reference training step, sharded) and config 5 (encode path) at N.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ -z "$SKIP_TESTS" ]; then
timeout 300 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/pytest_gpu_n$N.log 2>&1; grep -E "^FAILED|^ERROR|passed|failed|^E  " gpurun_out/pytest_gpu_n$N.log | cut -c1-300 | head -12
fi
timeout 200 $TR --master-port 29521 tools/dist_check.py > gpurun_out/dist_check_n$N.log 2>&1; echo "dist_check exit $?" >> gpurun_out/dist_check_n$N.log
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/dist_check_n$N.log | tail -4 | cut -c1-600
timeout 200 $TR --master-port 29522 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "exit $?" >> gpurun_out/bench_n$N.err
VQB_NCCL_ALLREDUCE=1 timeout 200 $TR --master-port 29523 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err
for f in gpurun_out/bench_n$N.json gpurun_out/bench_n${N}_nccl.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', d['n_gpus'], 'ms/step', d['ms_per_step'], 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['clocks'])"; done
tail -2 gpurun_out/bench_n$N.err | cut -c1-300
timeout 300 $TR --master-port 29524 tools/train_step_c4.py > gpurun_out/c4_n$N.json 2> gpurun_out/c4_n$N.err; echo "c4 exit $?"; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/c4_n$N.err | tail -3 | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/c4_n$N.json')); print({k: v for k, v in d.items() if k != 'boundary_parity'})"
timeout 200 $TR --master-port 29525 bench.py --workload encode > gpurun_out/encode_n$N.json 2> gpurun_out/encode_n$N.err; echo "encode exit $?"; cut -c1-1200 gpurun_out/encode_n$N.json
