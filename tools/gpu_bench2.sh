#!/bin/bash
# GPU box: short bench under a few environment settings.  usage: tools/gpu_bench2.sh "ENV=1 ..." "ENV=2 ..."
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
i=0
for e in "$@"; do
  i=$((i+1))
  env $e python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_x$i.json 2> gpurun_out/bench_x$i.err || { echo "bench failed"; tail -20 gpurun_out/bench_x$i.err; }
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_x$i.json"))
print("[$e] value %.3e e2e %.3e ms/step %.0f launches %d kernel_ms/launch %.2f dev_ms/step %.0f frac %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["kernel_ms_per_launch"]*d["gpu_launches"]/d["steps"], d["roofline"]["frac"]), d["e2e"]["host_seconds_per_step"])
PY
done
