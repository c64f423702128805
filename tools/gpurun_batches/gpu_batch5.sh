#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2b5_topo.txt 2>&1
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q -k multi_rank > gpurun_out/r2b5_pytest.log 2>&1; tail -3 gpurun_out/r2b5_pytest.log
bash tools/gpurun_batches/gpu_batch4.sh "8 4" "p2p"
mv gpurun_out/r2b4_n8_p2p.json gpurun_out/r2b5_n8_p2p.json; mv gpurun_out/r2b4_n4_p2p.json gpurun_out/r2b5_n4_p2p.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu-baseline --dp-mode nccl > gpurun_out/r2b5_n8_nccl.json 2> gpurun_out/r2b5_n8_nccl.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b5_n8_nccl.json").read().strip().splitlines()[-1]); print("n8 nccl value", round(d["value"],1), "ms", round(d["ms_per_step"],3))
PY
