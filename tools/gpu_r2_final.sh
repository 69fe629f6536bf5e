#!/bin/bash
# GPU box, round 2, last call: parity suite and the bench lines of the final revision (c4 default, c3 and c5s with their
# reference gap samples).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2f_pytest_gpu.log; tail -3 $O/r2f_pytest_gpu.log
report() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print("[%s] value %.3e e2e %.3e s/step %.2f gaps/s %.0f frac %.4f launches %d" % (sys.argv[1], d["value"], d["e2e"]["value"], d["ms_per_step"] / 1e3, d["e2e"]["gaps_per_s"], d["roofline"]["frac"], d["gpu_launches"]))
    print("   ", d["e2e"]["seconds_per_step_breakdown"]); print("   ", d["e2e"]["host_seconds_per_step"]); print("    parity", d.get("parity_sample"), d["clocks"])
except Exception as e:
    print("[%s] no bench line: %s" % (sys.argv[1], e))
PY
}
python bench.py --gpus 1 --steps 1 --warmup 1 --no-cpu-baseline > $O/r2f_bench_c4_n1.json 2> $O/r2f_bench_c4_n1.err || tail -5 $O/r2f_bench_c4_n1.err
report c4 $O/r2f_bench_c4_n1.json
python bench.py --workload c3 --gpus 1 --steps 2 --warmup 1 --no-cpu-baseline > $O/r2f_bench_c3_n1.json 2> $O/r2f_bench_c3_n1.err || tail -5 $O/r2f_bench_c3_n1.err
report c3 $O/r2f_bench_c3_n1.json
timeout 300 python bench.py --workload c5s --gpus 1 --steps 1 --warmup 1 --no-cpu-baseline > $O/r2f_bench_c5s_n1.json 2> $O/r2f_bench_c5s_n1.err || tail -5 $O/r2f_bench_c5s_n1.err
report c5s $O/r2f_bench_c5s_n1.json
