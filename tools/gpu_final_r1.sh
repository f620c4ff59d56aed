#!/bin/bash
# Last GPU visit of round 1 (seconds of budget left): the tests added late in the round, the pipelined-search A/B, config 5.
mkdir -p gpurun_out
( timeout 45 python -m pytest tests/test_gpu_parity.py -q --tb=short -k "generic or config5" > gpurun_out/f_new_tests.log 2>&1; echo "exit $?" >> gpurun_out/f_new_tests.log ) &
( VQB_SEARCH_PIPE=1 timeout 45 python -m pytest tests/test_gpu_tensor_search.py -q --tb=short > gpurun_out/f_pipe_tests.log 2>&1; echo "exit $?" >> gpurun_out/f_pipe_tests.log ) &
wait
tail -12 gpurun_out/f_new_tests.log | cut -c1-220
tail -8 gpurun_out/f_pipe_tests.log | cut -c1-220
VQB_SWEEP_PIPE_AB=1 VQB_SWEEP_POINTS="256x64,1024x64" timeout 25 python tools/sweep_c3.py > gpurun_out/f_sweep_pipe_ab.jsonl 2> gpurun_out/f_sweep_pipe_ab.err
python -c "
import json
for l in open('gpurun_out/f_sweep_pipe_ab.jsonl'):
    d = json.loads(l); print(d['K'], d['D'], 'fwd_ms', round(d['fwd_ms'], 4), 'pipe', round(d.get('fwd_ms_search_pipe', -1), 4))
"; tail -3 gpurun_out/f_sweep_pipe_ab.err | cut -c1-300
timeout 20 python tools/encode_c5.py > gpurun_out/f_encode_c5.json 2> gpurun_out/f_encode_c5.err
cut -c1-900 gpurun_out/f_encode_c5.json; tail -3 gpurun_out/f_encode_c5.err | cut -c1-300
