#!/bin/bash
# First GPU visit of round 2: what round 1 could no longer run.
#   1. the whole GPU suite (includes the tests added after the last full run of round 1)
#   2. the experimental variants written blind at the end of round 1 (three-slot p_code forward, column-split second epilogue
#      warpgroup, cluster-of-2 multicast codebook stream): correctness first, each under its own timeout (their mbarrier waits are
#      bounded: a protocol bug traps after 4 s instead of hanging), then the A/B timings of the config-3 sweep
#   3. bench + config-5 tool
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-300 | head -20
VQB_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_tensor_search.py -q --tb=short -k "column_split" > gpurun_out/exp_cs2.log 2>&1; echo "exit $?" >> gpurun_out/exp_cs2.log
VQB_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_tensor_search.py -q --tb=short -k "multicast_pair" > gpurun_out/exp_mc2.log 2>&1; echo "exit $?" >> gpurun_out/exp_mc2.log
tail -5 gpurun_out/exp_cs2.log | cut -c1-300; tail -5 gpurun_out/exp_mc2.log | cut -c1-300
# the three-slot p_code forward: the module parity tests and the bench with the variant switched on by the environment
VQB_FWD_X3=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q --tb=short -k "module_vs_reference or config2 or config5 or no_grad" > gpurun_out/exp_x3.log 2>&1; echo "exit $?" >> gpurun_out/exp_x3.log
tail -5 gpurun_out/exp_x3.log | cut -c1-300
VQB_FWD_X3=1 timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_x3.json 2> gpurun_out/bench_x3.err; cut -c1-400 gpurun_out/bench_x3.json
VQB_SWEEP_PIPE_AB=1 VQB_SWEEP_CS2_AB=1 timeout 300 python tools/sweep_c3.py > gpurun_out/sweep_ab_cs2.jsonl 2> gpurun_out/sweep_ab_cs2.err
VQB_SWEEP_MC2_AB=1 VQB_SWEEP_POINTS="4096x256,8192x256" timeout 300 python tools/sweep_c3.py > gpurun_out/sweep_ab_mc2.jsonl 2> gpurun_out/sweep_ab_mc2.err
# one-pass (1xTF32 + re-rank window) instead of three-pass search below K = 1024: search tests, then the sweep rows it changes
VQB_SEARCH_MODE=1 timeout 120 python -m pytest tests/test_gpu_tensor_search.py -q --tb=short > gpurun_out/exp_mode1.log 2>&1; echo "exit $?" >> gpurun_out/exp_mode1.log
tail -3 gpurun_out/exp_mode1.log | cut -c1-300
VQB_SEARCH_MODE=1 VQB_SWEEP_POINTS="256x64,1024x64" timeout 120 python tools/sweep_c3.py > gpurun_out/sweep_mode1.jsonl 2> gpurun_out/sweep_mode1.err
python - <<'PY'
import json
for f in ("gpurun_out/sweep_ab_cs2.jsonl", "gpurun_out/sweep_ab_mc2.jsonl", "gpurun_out/sweep_mode1.jsonl"):
    for l in open(f):
        d = json.loads(l)
        print(f[-12:], d["K"], d["D"], {k: round(v, 4) for k, v in d.items() if k.startswith("fwd_ms")})
PY
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
timeout 120 python tools/encode_c5.py > gpurun_out/encode_c5_n1.json 2> gpurun_out/encode_c5.err; cut -c1-400 gpurun_out/encode_c5_n1.json
