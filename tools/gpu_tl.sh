#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/timeline_bwd.py > gpurun_out/timeline_bwd.txt 2>&1
cat gpurun_out/timeline_bwd.txt
