#!/bin/bash
# GPU box, round 2, third call: A/B of the gather / unit-lookup variants on the C4 bench, engine parity tests on the combined build.
#   usage: tools/gpu_r2_ab3.sh <variant>...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
cp figbird_b200/_build/libfigbird_b200.so /tmp/base.so
report() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    ps = d.get("parity_sample") or {}
    print("[%s] value %.3e s/step %.2f kernels %.3f frac %.4f parity %s/%s of %s" % (sys.argv[1], d["value"], d["ms_per_step"] / 1e3,
          d["e2e"]["seconds_per_step_breakdown"]["kernels (busiest GPU)"], d["roofline"]["frac"], ps.get("identical_partial"), ps.get("identical_unmapped"), ps.get("gaps_sampled")))
except Exception as e:
    print("[%s] no bench line: %s" % (sys.argv[1], e))
PY
}
for v in base "$@"; do
  if [ "$v" = base ]; then cp /tmp/base.so figbird_b200/_build/libfigbird_b200.so; else cp figbird_b200/_variants/$v/libfigbird_b200.so figbird_b200/_build/libfigbird_b200.so; fi
  python bench.py --gpus 1 --steps 1 --warmup 0 --no-cpu-baseline > $O/r2c_ab_$v.json 2> $O/r2c_ab_$v.err || { echo "bench failed ($v)"; tail -5 $O/r2c_ab_$v.err; continue; }
  report $v $O/r2c_ab_$v.json
done
# the build with everything switched on through the engine parity tests (CUDA engine against the CPU oracle, item by item)
last="${@: -1}"
cp figbird_b200/_variants/$last/libfigbird_b200.so figbird_b200/_build/libfigbird_b200.so
timeout 600 python -m pytest tests/test_engine_parity_gpu.py -x -q > $O/r2c_pytest_engine_$last.log 2>&1; echo "engine parity ($last) rc=$?"; tail -2 $O/r2c_pytest_engine_$last.log
cp /tmp/base.so figbird_b200/_build/libfigbird_b200.so
