#!/bin/bash
# Short GPU visit: parity tests, timelines, bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 120 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 200 python tools/timeline_bwd.py > gpurun_out/timeline_bwd.txt 2>&1
timeout 200 python tools/timeline_fwd.py > gpurun_out/timeline_fwd.txt 2>&1
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
tail -15 gpurun_out/pytest_gpu.log; cat gpurun_out/timeline_bwd.txt | head -50; cut -c1-900 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
