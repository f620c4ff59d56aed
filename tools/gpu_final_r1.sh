#!/bin/bash
# Last GPU visits of round 1 (seconds of budget left).  Visit 1 ran: the tests added late in the round, the search tests with
# the pipelined x_lo forced on, a sweep A/B, config 5 (outputs: profiles/r1g_*).  Visit 2 (this form): the search tests with the
# final default (pipelined x_lo for <= 2 chunks per tile), then the sweep A/B with the guard in place.
mkdir -p gpurun_out
timeout 20 python -m pytest tests/test_gpu_tensor_search.py -q --tb=short > gpurun_out/g_search_tests.log 2>&1; echo "exit $?" >> gpurun_out/g_search_tests.log
tail -6 gpurun_out/g_search_tests.log | cut -c1-220
VQB_SWEEP_PIPE_AB=1 VQB_SWEEP_POINTS="256x64,1024x64" timeout 20 python tools/sweep_c3.py > gpurun_out/g_sweep_pipe_ab.jsonl 2> gpurun_out/g_sweep_pipe_ab.err
python -c "
import json
for l in open('gpurun_out/g_sweep_pipe_ab.jsonl'):
    d = json.loads(l); print(d['K'], d['D'], 'fwd_ms', round(d['fwd_ms'], 4), 'nopipe', round(d.get('fwd_ms_search_nopipe', -1), 4))
"; tail -3 gpurun_out/g_sweep_pipe_ab.err | cut -c1-300
