#!/bin/bash
# Not a test: DP overhead experiments on N GPUs (run on the GPU box from the repo root).
N=${N:-2}
run() { echo "== $*"; env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 2>gpurun_out/dp.err | tail -n 1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],3))"; }
run A=1
run MMEMO_BUCKET_MB=8
run MMEMO_SM_RESERVE=0
run NCCL_MAX_CTAS=8
run NCCL_MAX_CTAS=24
