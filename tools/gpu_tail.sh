#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
python tools/timeline_tail.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tee gpurun_out/timeline_tail_n1.txt
timeout 200 $TR --master-port 29531 tools/timeline_tail.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tee gpurun_out/timeline_tail_n$N.txt
bash tools/gpu_n.sh $N
timeout 300 python -m pytest tests -m gpu -q --timeout 280 2>&1 | tail -4
timeout 200 python bench.py --steps 200 --warmup 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1', d['ms_per_step'], d['value'], d['roofline']['kernel_ms'])"
