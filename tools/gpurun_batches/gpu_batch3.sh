#!/bin/bash
# round-2 batch 3 (1 GPU): A/B of tile-backward variants + row-extent tile mask; tests
mkdir -p gpurun_out
L=photo-slam_b200/lib
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2b3_$name.json 2> gpurun_out/r2b3_$name.err
  echo "$name rc=$?"
}
run default
run exact PSB_LIB=$PWD/$L/libpsb200_exact.so
run bwdppt4 PSB_BWD_PPT=4
run bwdppt1 PSB_BWD_PPT=1
run fwdppt4 PSB_FWD_PPT=4
run fwdppt1 PSB_FWD_PPT=1
run bwdmb4 PSB_LIB=$PWD/$L/libpsb200_bwdmb4.so
run bwdmb5 PSB_LIB=$PWD/$L/libpsb200_bwdmb5.so
run bwdmb8 PSB_LIB=$PWD/$L/libpsb200_bwdmb8.so
run default_B --config B
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b3_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2b3_")[1][:-5].ljust(10), "ms", round(d["ms_per_step"],3), "loss", d["loss"], {k: round(v["ms"],3) for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "unparsed", e)
PY
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -x -k "trainer or parity or dp" > gpurun_out/r2b3_pytest.log 2>&1; tail -3 gpurun_out/r2b3_pytest.log
