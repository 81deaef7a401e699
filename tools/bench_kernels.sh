#!/bin/bash
# Not a test: one bench run + per-kernel table (run on the GPU box).  Usage: tools/bench_kernels.sh TAG
TAG=${1:-x}
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || tail -c 800 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "ksum", round(d["roofline"]["kernel_time_sum_ms"],3), "launches", d["launches_per_step"])
for k in d.get("kernels", []):
    print(f'{k["kernel"][:64]:64s} {k["launches"]:3d} {k["us_per_launch"]:7.1f} us  {k["frac"]:.3f} {k["bound"]}')
PY
