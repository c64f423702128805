#!/bin/bash
# round-2 batch 1 (2 GPUs): full -m gpu suite, bench N=1, N=2 (p2p and nccl)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2b1_gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/r2b1_gpus.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -rs --durations=15 > gpurun_out/r2b1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b1_pytest.log
tail -5 gpurun_out/r2b1_pytest.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2b1_bench_n1.json 2> gpurun_out/r2b1_bench_n1.err
echo "bench n1 rc=$?"
for mode in p2p nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --dp-mode $mode > gpurun_out/r2b1_bench_n2_$mode.json 2> gpurun_out/r2b1_bench_n2_$mode.err
  echo "bench n2 $mode rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b1_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["config"].get("parallelism","")[:60])
    except Exception as e:
        print(f, "unparsed", e)
PY
