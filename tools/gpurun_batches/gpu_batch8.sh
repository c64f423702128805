#!/bin/bash
mkdir -p gpurun_out
n=8
for G in 4 1 2; do
  PSB_DP_GROUPS=$G timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline --dp-mode p2p > gpurun_out/r2b8_n${n}_g$G.json 2> gpurun_out/r2b8_n${n}_g$G.err
  echo "bench n=$n groups=$G rc=$?"
done
PSB_DP_GROUPS=4 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 50 --warmup 5 --no-cpu-baseline --dp-mode p2p > gpurun_out/r2b8_n4_g4.json 2> gpurun_out/r2b8_n4_g4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b8_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2b8_")[1][:-5].ljust(10), "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in d.get("dp_stages_ms_rank0",{}).items()})
    except Exception as e:
        print(f, "unparsed", e)
PY
