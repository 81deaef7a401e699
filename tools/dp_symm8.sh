#!/bin/bash
# Not a test: 8-GPU scaling check of both gradient transports (run on the GPU box).
N=${N:-8}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/test_allreduce.py > gpurun_out/ar_$N.log 2>&1
echo "exit $?"; grep "nvls\|peer\|nccl\|multicast" gpurun_out/ar_$N.log | grep -v "blocks  4\|blocks 32\|33.6" | tail -n 40
run() { echo "== $*"; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 2>gpurun_out/dp.err | tail -n 1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],3))" || tail -n 15 gpurun_out/dp.err; }
run MMEMO_DP_TRANSPORT=symm
run MMEMO_DP_TRANSPORT=nccl
run MMEMO_DP_TRANSPORT=symm MMEMO_BUCKET_MB=16
timeout 200 python bench.py --steps 30 --warmup 3 2>/dev/null | tail -n 1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],3))"
