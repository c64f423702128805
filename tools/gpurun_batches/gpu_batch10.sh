#!/bin/bash
mkdir -p gpurun_out
PSB_DP_TIMEOUT_MS=8000 timeout 900 python -m pytest tests/test_dp_gpu.py tests/test_trainer_gpu.py -m gpu -q -rs > gpurun_out/r2b10_pytest.log 2>&1; tail -3 gpurun_out/r2b10_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --dp-mode p2p > gpurun_out/r2b10_n2.json 2> gpurun_out/r2b10_n2.err; echo "n2 rc=$?"
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2b10_n1.json 2> gpurun_out/r2b10_n1.err; echo "n1 rc=$?"
timeout 600 python tools/densify_bench.py 500000 3000000 > gpurun_out/r2b10_densify.jsonl 2> gpurun_out/r2b10_densify.err; echo "densify rc=$?"; cat gpurun_out/r2b10_densify.jsonl | cut -c1-400
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b10_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2b10_")[1][:-5].ljust(6), "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in d.get("dp_stages_ms_rank0",{}).items()}, {k: round(v["ms"],3) for k,v in d.get("stages",{}).items()}, d["clocks"].get("samples"))
    except Exception as e:
        print(f, "unparsed", e)
PY
