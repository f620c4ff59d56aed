#!/bin/bash
# Round 2, first GPU visit: the third-generation parity-mode kernels (vqb_fwd_pcode_kernel / vqb_bwd_pcode_kernel) against the
# second-generation ones on the same box: parity first (each combination in its own process: a trapped kernel kills the CUDA
# context), then timelines, the bench A/B, the launch list and the ncu --set full captures.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
SUB="module_vs_reference or tensor_core_backward or fused_backward_tail or config2 or no_grad or fused_mode"
run() { name=$1; shift; env "$@" timeout 300 python -m pytest tests/test_gpu_parity.py -q --tb=short -x -k "$SUB" > gpurun_out/par_$name.log 2>&1; echo "exit $?" >> gpurun_out/par_$name.log; echo "== $name: $(grep -E 'passed|failed|error' gpurun_out/par_$name.log | tail -1)"; }
run new_new X=1
run new_fwd_old_bwd VQB_BWD_KERNEL=h2
run old_fwd_new_bwd VQB_FWD_OLD=1
grep -E "^FAILED|^ERROR|Error|error:" gpurun_out/par_new_new.log | head -10 | cut -c1-300
timeout 200 python tools/timeline_pc.py > gpurun_out/timeline_pc.txt 2>&1; head -60 gpurun_out/timeline_pc.txt
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-1600 gpurun_out/bench.json
VQB_FWD_OLD=1 VQB_BWD_KERNEL=h2 timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_old.json 2> gpurun_out/bench_old.err; cut -c1-400 gpurun_out/bench_old.json
python - <<'PY'
import json
for f in ("bench", "bench_old"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "ms/step %.4f" % d["ms_per_step"], d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"], "e2e %.3g" % d["e2e"]["value"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-300 | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
bash tools/ncu_full.sh vqb_bwd_pcode_kernel bwd
bash tools/ncu_full.sh vqb_fwd_pcode_kernel fwd
ls -la gpurun_out/*.ncu-rep
