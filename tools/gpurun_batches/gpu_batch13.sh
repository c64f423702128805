#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --maxfail=5 -rs > gpurun_out/r2b13_pytest.log 2>&1; tail -4 gpurun_out/r2b13_pytest.log
timeout 300 python bench.py --steps 100 --warmup 5 > gpurun_out/r2b13_bench_D.json 2> gpurun_out/r2b13_bench_D.err; echo "bench D rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b13_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b13_ncu_launch.log 2>&1; echo "ncu launch rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on --launch-skip 57 --launch-count 22 -o gpurun_out/r2b13_step python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b13_ncu_full.log 2>&1; echo "ncu full rc=$?"
PSB_DENSIFY_BREAKDOWN=1 timeout 200 python tools/densify_bench.py 3000000 > gpurun_out/r2b13_densify_breakdown.jsonl 2> gpurun_out/r2b13_densify_breakdown.err; tail -3 gpurun_out/r2b13_densify_breakdown.err
timeout 200 python tools/densify_bench.py 500000 3000000 > gpurun_out/r2b13_densify.jsonl 2> gpurun_out/r2b13_densify.err; cut -c1-330 gpurun_out/r2b13_densify.jsonl
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b13_bench_D.json").read().strip().splitlines()[-1])
print("D value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k: (round(v["ms"],3), round(v["frac_of_hbm_peak"],2), v.get("issue_frac") and round(v["issue_frac"],2)) for k,v in d["stages"].items()}, d["clocks"], d["config"].get("optimizer_live_rows"))
print(json.dumps(d["roofline"])[:900])
PY
