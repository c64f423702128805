#!/bin/bash
# N-GPU data-parallel bench: usage gpu_batch4.sh "2 4 8" [modes]
mkdir -p gpurun_out
NS=${1:-2}
MODES=${2:-p2p}
for n in $NS; do for mode in $MODES; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline --dp-mode $mode > gpurun_out/r2b4_n${n}_$mode.json 2> gpurun_out/r2b4_n${n}_$mode.err
  echo "bench n=$n $mode rc=$?"
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b4_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2b4_")[1][:-5].ljust(10), "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in d.get("dp_stages_ms_rank0",{}).items()})
    except Exception as e:
        print(f, "unparsed", e)
PY
