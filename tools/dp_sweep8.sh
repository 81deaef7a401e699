#!/bin/bash
# Not a test: 8-GPU DP experiments (run on the GPU box from the repo root).
N=${N:-8}
run() { echo "== $*"; env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 2>gpurun_out/n8.err | tail -n 1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"; }
run NCCL_MAX_CTAS=32 MMEMO_BUCKET_MB=17
run NCCL_MAX_CTAS=32 MMEMO_BUCKET_MB=64
run MMEMO_DP_NO_COMM=1
NCCL_MAX_CTAS=32 timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/nccl_bench.py 2>/dev/null | grep allreduce
