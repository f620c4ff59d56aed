#!/bin/bash
# ncu --set full capture of the forward and backward kernels of the bench workload (one GPU).
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s 6 -c 2 -f -o gpurun_out/prof_$2 \
    python bench.py --steps 4 --warmup 3 --no-sweep > gpurun_out/ncu_full_$2.log 2>&1
tail -3 gpurun_out/ncu_full_$2.log
