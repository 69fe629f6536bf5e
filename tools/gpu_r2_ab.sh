#!/bin/bash
# GPU box: A/B of kernel build variants on the C4 bench (N=1).  usage: tools/gpu_r2_ab.sh <variant>...   ("base" = product build)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp figbird_b200/_build/libfigbird_b200.so /tmp/base.so
for v in "$@"; do
  if [ "$v" = base ]; then cp /tmp/base.so figbird_b200/_build/libfigbird_b200.so; else cp figbird_b200/_variants/$v/libfigbird_b200.so figbird_b200/_build/libfigbird_b200.so; fi
  python bench.py --gpus 1 --steps ${STEPS:-1} --warmup 1 --no-cpu-baseline ${WL:+--workload $WL} > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err || { echo "bench failed ($v)"; tail -20 gpurun_out/bench_ab_$v.err; continue; }
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_ab_$v.json"))
print("[$v] value %.3e e2e %.3e s/step %.2f kernels %.2f frac %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]/1e3, d["e2e"]["seconds_per_step_breakdown"]["kernels (busiest GPU)"], d["roofline"]["frac"]))
PY
done
cp /tmp/base.so figbird_b200/_build/libfigbird_b200.so
