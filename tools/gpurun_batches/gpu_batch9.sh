#!/bin/bash
mkdir -p gpurun_out
n=8
for rot in 1 0; do
  PSB_DP_ROTATE=$rot timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline --dp-mode p2p > gpurun_out/r2b9_n${n}_rot$rot.json 2> gpurun_out/r2b9_n${n}_rot$rot.err
  echo "bench n=$n rotate=$rot rc=$?"
done
PSB_DP_TIMEOUT_MS=8000 timeout 300 python -m pytest tests/test_dp_gpu.py -m gpu -q -k multi_rank > gpurun_out/r2b9_pytest.log 2>&1; tail -2 gpurun_out/r2b9_pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b9_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2b9_")[1][:-5].ljust(10), "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in d.get("dp_stages_ms_rank0",{}).items() if k in ("push_backward","wait_grads","shard_adam","shard_adam_frest_part","wait_params")})
    except Exception as e:
        print(f, "unparsed", e)
PY
