#!/bin/bash
mkdir -p gpurun_out
n=${1:-2}
PSB_DP_TIMEOUT_MS=8000 timeout 900 python -m pytest tests/test_dp_gpu.py tests/test_quality_gpu.py -m gpu -q -rs -s > gpurun_out/r2b7_pytest.log 2>&1; grep -E "passed|failed|PSNR over|densify @|DP_WORKER" gpurun_out/r2b7_pytest.log | head
for G in 4 1 2 8; do
  PSB_DP_GROUPS=$G timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline --dp-mode p2p > gpurun_out/r2b7_n${n}_g$G.json 2> gpurun_out/r2b7_n${n}_g$G.err
  echo "bench n=$n groups=$G rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b7_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2b7_")[1][:-5].ljust(10), "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in d.get("dp_stages_ms_rank0",{}).items()})
    except Exception as e:
        print(f, "unparsed", e)
PY
