#!/bin/bash
# round-2 batch 2a (1 GPU): full -m gpu suite, bench N=1 (configs D, B, C), ncu launch list + full capture of one iteration
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -rs --durations=10 > gpurun_out/r2b2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b2_pytest.log
tail -4 gpurun_out/r2b2_pytest.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2b2_bench_D.json 2> gpurun_out/r2b2_bench_D.err; echo "bench D rc=$?"
timeout 600 python bench.py --config B --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2b2_bench_B.json 2> gpurun_out/r2b2_bench_B.err; echo "bench B rc=$?"
timeout 600 python bench.py --config C --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2b2_bench_C.json 2> gpurun_out/r2b2_bench_C.err; echo "bench C rc=$?"
# launch list of the bench command (cold-cache, serialised): shares of the step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b2_ncu_launch.log 2>&1; echo "ncu launch rc=$?"
# full capture of ONE training iteration (skip the warm-up launches: 3 warm-up steps x ~21 launches + memsets are not kernels)
timeout 1200 ncu --set full --clock-control none --import-source on --launch-skip 63 --launch-count 21 -o gpurun_out/r2b2_step python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b2_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/r2b2_step.ncu-rep
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b2_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "render Mpix/s", round(d["render_mpix_per_s"],1))
        print("   ", {k: round(v["ms"],3) for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "unparsed", e)
PY
