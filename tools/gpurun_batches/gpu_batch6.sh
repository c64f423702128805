#!/bin/bash
# validation + profiles (1 GPU): full -m gpu suite, bench D/B/C, ncu launch list + full capture of one iteration
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -rs -s > gpurun_out/r2b6_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b6_pytest.log
grep -E "passed|failed|SKIPPED|PSNR|densify @|distCUDA2 over|FAILED" gpurun_out/r2b6_pytest.log | head -30
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r2b6_bench_D.json 2> gpurun_out/r2b6_bench_D.err; echo "bench D rc=$?"
timeout 600 python bench.py --config B --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2b6_bench_B.json 2> gpurun_out/r2b6_bench_B.err; echo "bench B rc=$?"
timeout 600 python bench.py --config C --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2b6_bench_C.json 2> gpurun_out/r2b6_bench_C.err; echo "bench C rc=$?"
timeout 900 python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/r2b6_bench_ref_D.json 2> gpurun_out/r2b6_bench_ref_D.err; echo "ref D rc=$?"
timeout 900 python bench.py --impl reference --config B --steps 30 --warmup 5 > gpurun_out/r2b6_bench_ref_B.json 2> gpurun_out/r2b6_bench_ref_B.err; echo "ref B rc=$?"
timeout 900 python bench.py --impl reference --config C --steps 30 --warmup 5 > gpurun_out/r2b6_bench_ref_C.json 2> gpurun_out/r2b6_bench_ref_C.err; echo "ref C rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b6_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b6_ncu_launch.log 2>&1; echo "ncu launch rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on --launch-skip 57 --launch-count 22 -o gpurun_out/r2b6_step python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b6_ncu_full.log 2>&1; echo "ncu full rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b6_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2b6_bench_")[1][:-5].ljust(8), "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "Mpix/s", round(d["render_mpix_per_s"],1))
        if "stages" in d: print("    ", {k: round(v["ms"],3) for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "unparsed", e)
PY
