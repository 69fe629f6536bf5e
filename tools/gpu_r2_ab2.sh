#!/bin/bash
# GPU box, round 2, second call: parity suite on the current host code, A/B of kernel build variants on the C4 bench (one timed
# step each, kernels only matter), the two-library workload c3, and a `--set full` capture aimed at the launches that dominate
# C4 (fb_em_kernel<true, 512>, unmapped mode).   usage: tools/gpu_r2_ab2.sh <variant>...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2b_pytest_gpu.log; tail -3 $O/r2b_pytest_gpu.log
cp figbird_b200/_build/libfigbird_b200.so /tmp/base.so
report() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print("[%s] value %.3e e2e %.3e s/step %.2f kernels %.2f frac %.4f idle %.2f serial %.2f parity %s" % (sys.argv[1], d["value"], d["e2e"]["value"], d["ms_per_step"] / 1e3,
          d["e2e"]["seconds_per_step_breakdown"]["kernels (busiest GPU)"], d["roofline"]["frac"],
          list(d["e2e"]["seconds_per_step_breakdown"].values())[1], list(d["e2e"]["seconds_per_step_breakdown"].values())[2], d.get("parity_sample")))
    print("   host", d["e2e"]["host_seconds_per_step"])
except Exception as e:
    print("[%s] no bench line: %s" % (sys.argv[1], e))
PY
}
for v in warm base "$@"; do
  if [ "$v" = base ] || [ "$v" = warm ]; then cp /tmp/base.so figbird_b200/_build/libfigbird_b200.so; else cp figbird_b200/_variants/$v/libfigbird_b200.so figbird_b200/_build/libfigbird_b200.so; fi
  python bench.py --gpus 1 --steps 1 --warmup 0 --no-cpu-baseline > $O/r2b_ab_$v.json 2> $O/r2b_ab_$v.err || { echo "bench failed ($v)"; tail -5 $O/r2b_ab_$v.err; continue; }
  report $v $O/r2b_ab_$v.json
done
cp /tmp/base.so figbird_b200/_build/libfigbird_b200.so
# two-library workload (BASELINE configs[2])
python bench.py --workload c3 --gpus 1 --steps 2 --warmup 1 --no-cpu-baseline > $O/r2b_bench_c3.json 2> $O/r2b_bench_c3.err || tail -5 $O/r2b_bench_c3.err
report c3 $O/r2b_bench_c3.json
# the launches that dominate C4: big-bucket kernel, unmapped mode (the partial-mode launches come first and are skipped)
python bench.py --workload c4s --gpus 1 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:fb_em_kernel.*512" --launch-skip 80 -c 8 -f -o $O/r2b_prof_big python bench.py --workload c4s --gpus 1 --steps 1 --warmup 0 --no-cpu-baseline > $O/r2b_ncu_big.log 2>&1
echo "full capture rc=$?"; tail -2 $O/r2b_ncu_big.log
ncu -i $O/r2b_prof_big.ncu-rep --page raw --csv > $O/r2b_fb_em_kernel_big_raw.csv 2>/dev/null
ls -la $O | grep r2b | tail -20
