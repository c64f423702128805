#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -rs > gpurun_out/r2b3b_pytest.log 2>&1; tail -6 gpurun_out/r2b3b_pytest.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2b3b_D.json 2> gpurun_out/r2b3b_D.err; echo "D rc=$?"
timeout 300 python bench.py --config B --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2b3b_B.json 2> gpurun_out/r2b3b_B.err; echo "B rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b3b_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2b3b_")[1][:-5].ljust(10), "ms", round(d["ms_per_step"],3), "loss", d["loss"], {k: round(v["ms"],3) for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "unparsed", e)
PY
