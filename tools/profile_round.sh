#!/bin/bash
# Not a test: everything profiles/ cites for a round, in one GPU-box call (run from the repo root):
#   bench lines (ours + reference arm), NVTX-renamed ncu launch lists of one step of cfg2 / cfg1a /
#   cfg3c / cfg4 (durations + DRAM bytes per launch), and one `--set full` capture of the cfg-2
#   step's hot kernels.  Each ncu pass runs only after the same command exited 0 without ncu.
set -u
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err || tail -n 5 gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_n1_reference.json 2> gpurun_out/r02_ref.err || tail -n 5 gpurun_out/r02_ref.err
for W in cfg2 cfg1a cfg3c cfg4; do
  python bench.py --ncu-step $W > /dev/null 2> gpurun_out/ncu_step_$W.err || { echo "plain $W failed"; tail -n 5 gpurun_out/ncu_step_$W.err; continue; }
  timeout 600 ncu --nvtx --print-nvtx-rename kernel --profile-from-start off --clock-control none \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --csv --log-file gpurun_out/r02_ncu_${W}_families.csv python bench.py --ncu-step $W > gpurun_out/ncu_$W.log 2>&1
  echo "ncu $W exit $?"
done
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'gemm_tc_kernel|resattn_bwd_tc_kernel|resattn_fwd_tc_kernel|ln_bwd_fused_vec|ln_fwd_vec' \
  --launch-skip 20 --launch-count 36 -f -o gpurun_out/r02_full_cfg2 python bench.py --ncu-step cfg2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep
