#!/bin/bash
mkdir -p gpurun_out
PSB_DP_TIMEOUT_MS=8000 timeout 400 python -m pytest tests/test_dp_gpu.py -m gpu -q -x > gpurun_out/r2b12_pytest.log 2>&1; tail -3 gpurun_out/r2b12_pytest.log
for b in 1 0; do
PSB_DP_BULK=$b timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline --dp-mode p2p > gpurun_out/r2b12_n2_bulk$b.json 2> gpurun_out/r2b12_n2_bulk$b.err; echo "n2 bulk=$b rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b12_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2b12_")[1][:-5].ljust(9), "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in d.get("dp_stages_ms_rank0",{}).items() if k in ("push_backward","shard_adam","shard_adam_frest_part")})
    except Exception as e:
        print(f, "unparsed", e)
PY
