#!/bin/bash
# Round 2 GPU visit: PDL A/B of the step, near-tie test, the search-kernel experiments left from round 1 (CS2 / MC2 / one-pass below K=1024)
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 60 -k "near_ties or realistic or module_vs" > gpurun_out/pytest_a.log 2>&1; tail -3 gpurun_out/pytest_a.log | cut -c1-300
b() { name=$1; shift; env "$@" timeout 120 python bench.py --steps 200 --warmup 10 --no-sweep > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$name.json')); print('$name', 'ms/step %.4f' % d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e %.3g' % d['e2e']['value'])"; }
b pdl_all X=1
b no_bwd_pdl VQB_BWD_NO_PDL=1
b no_asm_pdl VQB_ASM_NO_PDL=1
b no_both VQB_BWD_NO_PDL=1 VQB_ASM_NO_PDL=1
VQB_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_tensor_search.py -q --tb=short --timeout 60 -k "column_split" > gpurun_out/exp_cs2.log 2>&1; echo "exit $?" >> gpurun_out/exp_cs2.log
VQB_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_tensor_search.py -q --tb=short --timeout 60 -k "multicast_pair" > gpurun_out/exp_mc2.log 2>&1; echo "exit $?" >> gpurun_out/exp_mc2.log
tail -4 gpurun_out/exp_cs2.log | cut -c1-300; tail -4 gpurun_out/exp_mc2.log | cut -c1-300
VQB_SWEEP_CS2_AB=1 timeout 200 python tools/sweep_c3.py > gpurun_out/sweep_ab_cs2.jsonl 2> gpurun_out/sweep_ab_cs2.err
VQB_SWEEP_MC2_AB=1 VQB_SWEEP_POINTS="4096x256,8192x256" timeout 120 python tools/sweep_c3.py > gpurun_out/sweep_ab_mc2.jsonl 2> gpurun_out/sweep_ab_mc2.err
VQB_SEARCH_MODE=1 timeout 100 python -m pytest tests/test_gpu_tensor_search.py -q --tb=short --timeout 60 > gpurun_out/exp_mode1.log 2>&1; echo "exit $?" >> gpurun_out/exp_mode1.log; tail -3 gpurun_out/exp_mode1.log | cut -c1-300
VQB_SEARCH_MODE=1 VQB_SWEEP_POINTS="256x64,1024x64" timeout 100 python tools/sweep_c3.py > gpurun_out/sweep_mode1.jsonl 2> gpurun_out/sweep_mode1.err
python - <<'PY'
import json
for f in ("gpurun_out/sweep_ab_cs2.jsonl", "gpurun_out/sweep_ab_mc2.jsonl", "gpurun_out/sweep_mode1.jsonl"):
    try:
        for l in open(f):
            d = json.loads(l)
            print(f[-12:], d["K"], d["D"], {k: round(v, 4) for k, v in d.items() if k.startswith("fwd_ms")})
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -2 gpurun_out/sweep_ab_cs2.err gpurun_out/sweep_ab_mc2.err | cut -c1-300
