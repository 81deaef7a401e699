#!/bin/bash
# Not a test: does reserving SMs for the symmetric-memory all-reduce (MMEMO_SYMM_SM_RESERVE) pay?
# Run on the GPU box from the repo root with N GPUs.
N=${N:-2}
run() { echo "== $*"; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 --configs none --no-cpu-baseline 2>gpurun_out/dp.err | tail -n 1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); dp=d.get('dp') or {}; print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],4), 'no_comm', dp.get('ms_per_step_without_comm'), 'exposed', dp.get('exposed_comm_us'))" || tail -n 15 gpurun_out/dp.err; }
run A=1
run MMEMO_SYMM_SM_RESERVE=16
run MMEMO_SYMM_SM_RESERVE=16 MMEMO_RESERVE_LAUNCHES=8
run MMEMO_SYMM_SM_RESERVE=32 MMEMO_COMM_BLOCKS=32
run MMEMO_BUCKET_MB=8
run MMEMO_BUCKET_MB=8 MMEMO_SYMM_SM_RESERVE=16
run MMEMO_BUCKET_MB=48
