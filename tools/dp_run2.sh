#!/bin/bash
# Not a test: 2-GPU data-parallel check + sweep (run on the GPU box from the repo root).
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/t.log 2>&1; tail -4 gpurun_out/t.log
python bench.py --steps 20 --warmup 5 --configs none --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N1', d['ms_per_step'], d['roofline']['frac'])"
MMEMO_DP_NO_COMM=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dp_profile.py > gpurun_out/dp_prof_n2.txt 2>&1
run() { echo "== $*"; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 --configs none --no-cpu-baseline 2>gpurun_out/dp.err | tail -n 1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); dp=d.get('dp') or {}; print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],4), 'no_comm', dp.get('ms_per_step_without_comm'), 'exposed', dp.get('exposed_comm_us'), dp.get('grad_check',{}).get('max_rel_err'), dp.get('bucket_bytes'))" || tail -n 15 gpurun_out/dp.err; }
run A=1
run MMEMO_SYMM_SM_RESERVE=0
run MMEMO_BUCKET_MB=8
run MMEMO_BUCKET_MB=24
