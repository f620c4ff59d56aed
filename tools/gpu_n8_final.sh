#!/bin/bash
# 8-GPU box, final code of the round: 1 GPU and 8 GPUs at the driver's arguments (--steps 20 --warmup 5) and at 300 steps.
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
show() { python -c "
import json,sys; d=json.load(open('$1')); print('$2', 'n', d['n_gpus'], 'ms/step %.4f' % d['ms_per_step'], 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['clocks'])"; }
mkdir -p gpurun_out
timeout 200 python bench.py --steps 20 --warmup 5 --no-sweep > gpurun_out/f_n1_20.json 2> gpurun_out/f_n1_20.err; show gpurun_out/f_n1_20.json n1_20
timeout 200 python bench.py --steps 300 --warmup 10 --no-sweep > gpurun_out/f_n1_300.json 2> gpurun_out/f_n1_300.err; show gpurun_out/f_n1_300.json n1_300
timeout 200 $TR --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/f_n8_20.json 2> gpurun_out/f_n8_20.err; show gpurun_out/f_n8_20.json n8_20
timeout 200 $TR --master-port 29523 bench.py --gpus $N --steps 300 --warmup 10 > gpurun_out/f_n8_300.json 2> gpurun_out/f_n8_300.err; show gpurun_out/f_n8_300.json n8_300
