#!/bin/bash
# GPU box: A/B of kernel build variants.  Variants are built beforehand into figbird_b200/_variants/<name>/ with
#   make -C figbird_b200/csrc OUT=../_variants/<name> EXTRA="-DFB_X=1"
# usage: tools/gpu_variants.sh [-t] <name>...   ("base" = the product build; -t: also run the engine parity tests per variant)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tests=0; if [ "$1" = "-t" ]; then tests=1; shift; fi
cp figbird_b200/_build/libfigbird_b200.so /tmp/base.so
for v in "$@"; do
  if [ "$v" = base ]; then cp /tmp/base.so figbird_b200/_build/libfigbird_b200.so; else cp figbird_b200/_variants/$v/libfigbird_b200.so figbird_b200/_build/libfigbird_b200.so; fi
  if [ $tests = 1 ]; then python -m pytest tests/test_engine_parity_gpu.py -x -q 2>&1 | tail -2; fi
  FIGBIRD_PHASES=1 FIGBIRD_MODEL_TIMING=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_var_$v.json 2> gpurun_out/bench_var_$v.err || { echo "bench failed ($v)"; tail -20 gpurun_out/bench_var_$v.err; continue; }
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_var_$v.json"))
print("[$v] value %.3e e2e %.3e ms/step %.0f dev_ms/step %.0f frac %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"]*d["gpu_launches"]/d["steps"], d["roofline"]["frac"]))
PY
  grep "^phases" gpurun_out/bench_var_$v.err | tail -1; grep "^learnModel" gpurun_out/bench_var_$v.err | tail -8 | tr "\n" " "; echo
done
cp /tmp/base.so figbird_b200/_build/libfigbird_b200.so
