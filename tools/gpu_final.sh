#!/bin/bash
# Last 1-GPU visit of the round: the whole GPU suite, both bench arms at the driver's own arguments and at 200 steps, ncu --set full
# of the streamed search (K = 8192, D = 256 and D = 64).
mkdir -p gpurun_out; rm -f gpurun_out/test_records.jsonl
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|Error" gpurun_out/pytest_gpu.log | cut -c1-300 | head -10
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
timeout 400 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_driver_args.json 2> gpurun_out/bench_driver_args.err; echo "bench(driver args) exit $?"
python - <<'PY'
import json
for f in ("bench", "bench_driver_args"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "ms/step %.4f" % d["ms_per_step"], "value %.4g" % d["value"], d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"], "e2e %.4g" % d["e2e"]["value"], d.get("clocks"))
        for p in d.get("sweep", {}).get("points", []):
            print("   ", p["K"], p["D"], "fwd_ms %.3f" % p["fwd_ms"], "tf32 frac %.3f" % p["tensor_frac_of_tf32_peak"], "scatter hbm %.3f" % p["scatter_hbm_frac"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
if [ -z "$NO_NCU" ]; then
for pt in 8192x256 8192x64; do
VQB_SWEEP_POINTS=$pt timeout 600 ncu --set full --clock-control none --import-source on -k regex:vqb_fwd_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_search_$pt \
    python tools/sweep_c3.py > gpurun_out/ncu_search_$pt.log 2>&1
tail -1 gpurun_out/ncu_search_$pt.log | cut -c1-200
done
fi
