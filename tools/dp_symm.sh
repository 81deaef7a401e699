#!/bin/bash
# Not a test: libmmemo symmetric-memory all-reduce vs NCCL on N GPUs (run on the GPU box).
N=${N:-2}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/test_allreduce.py > gpurun_out/ar_$N.log 2>&1
echo "exit $?"; grep -v "^W\|^\*\*\*" gpurun_out/ar_$N.log | tail -n 60
run() { echo "== $*"; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 2>gpurun_out/dp.err | tail -n 1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],3))" || tail -n 15 gpurun_out/dp.err; }
run MMEMO_DP_TRANSPORT=symm
run MMEMO_DP_TRANSPORT=symm MMEMO_COMM_BLOCKS=8
run MMEMO_DP_TRANSPORT=symm MMEMO_COMM_BLOCKS=4
run MMEMO_DP_TRANSPORT=symm MMEMO_COMM_BLOCKS=32
run MMEMO_DP_TRANSPORT=symm MMEMO_BUCKET_MB=4
run MMEMO_DP_TRANSPORT=symm MMEMO_BUCKET_MB=16
run MMEMO_DP_TRANSPORT=symm MMEMO_DP_NO_COMM=1
run MMEMO_DP_TRANSPORT=nccl
