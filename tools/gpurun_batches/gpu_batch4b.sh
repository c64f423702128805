#!/bin/bash
# A/B of the signalling strategy at N GPUs: usage gpu_batch4b.sh N
mkdir -p gpurun_out
n=${1:-2}
for sig in kernel fence; do
  PSB_DP_SIGNAL=$sig timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline --dp-mode p2p > gpurun_out/r2b4b_n${n}_$sig.json 2> gpurun_out/r2b4b_n${n}_$sig.err
  echo "bench n=$n $sig rc=$?"
done
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q -k multi_rank > gpurun_out/r2b4b_pytest.log 2>&1; tail -3 gpurun_out/r2b4b_pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b4b_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2b4b_")[1][:-5].ljust(10), "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in d.get("dp_stages_ms_rank0",{}).items()})
    except Exception as e:
        print(f, "unparsed", e)
PY
