#!/bin/bash
# Not a test: ncu captures for profiles/ (run on the GPU box, one mode per gpurun call).
#   tools/ncu_capture.sh list   -> per-launch durations of a 2-step bench run (CSV)
#   tools/ncu_capture.sh full   -> --set full of one step's worth of the hot kernels (.ncu-rep)
set -u
MODE=${1:-list}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || { echo "plain run failed"; tail -n 20 gpurun_out/ncu_plain.err; exit 1; }
tail -c 300 gpurun_out/ncu_plain.json; echo
if [ "$MODE" = list ]; then
  timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo "ncu exit $?"; wc -l gpurun_out/launches.csv
else
  timeout 1200 ncu --set full --clock-control none \
    -k regex:'gemm_tc_kernel|resattn_.*_tc_kernel|ln_bwd_fused_vec|ln_fwd_vec|rowsum_kernel|cast_multi' \
    --launch-skip 440 --launch-count 26 -o gpurun_out/r01_full $CMD --no-graph > gpurun_out/ncu_full.log 2>&1
  echo "ncu exit $?"; ls -la gpurun_out/r01_full.ncu-rep
fi
