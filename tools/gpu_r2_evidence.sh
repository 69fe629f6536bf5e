#!/bin/bash
# GPU box, round 2: parity suite, the C4 bench line (N=1), one diagnostic step (tick log + per-bucket kernel time), the ncu launch
# list of the bench command and one `--set full` capture of fb_em_kernel.  Everything lands in gpurun_out/r2_*.
#   usage: tools/gpu_r2_evidence.sh [skip_tests]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
{ nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv; nproc; free -g | head -2; } > $O/r2_box.txt 2>&1
if [ -z "$1" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2_pytest_gpu.log; tail -4 $O/r2_pytest_gpu.log
fi
# (1) the bench line, default workload (c4), with the in-line cpu_baseline
python bench.py --gpus 1 --steps 1 --warmup 1 > $O/r2_bench_c4_n1.json 2> $O/r2_bench_c4_n1.err || { echo "bench failed"; tail -30 $O/r2_bench_c4_n1.err; }
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_c4_n1.json"))
    print("c4 N=1 value %.3e e2e %.3e s/step %.2f gaps/s %.0f frac %.4f prep %.0fs launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"] / 1e3, d["e2e"]["gaps_per_s"], d["roofline"]["frac"], d["data_prep_s"], d["gpu_launches"]))
    print(d["e2e"]["seconds_per_step_breakdown"]); print(d["e2e"]["host_seconds_per_step"]); print(d["clocks"]); print("d2h MB/step", d["e2e"]["d2h_bytes_per_step"] / 1e6)
    print(d.get("cpu_baseline"))
except Exception as e:
    print("no bench line:", e)
PY
# (2) one diagnostic step: tick log per lane and kernel time per shared-memory bucket
FIGBIRD_TICK_LOG=1 FIGBIRD_BUCKET_STATS=1 python bench.py --gpus 1 --steps 1 --warmup 0 --no-cpu-baseline > $O/r2_diag_c4.json 2> $O/r2_diag_c4.err; grep -E "^(ticklog|bucket)" $O/r2_diag_c4.err | cut -c1-1500 | tail -12
# (3) launch list of the bench command (one step)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_c4.csv python bench.py --gpus 1 --steps 1 --warmup 0 --no-cpu-baseline > $O/r2_ncu_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < $O/r2_launches_c4.csv)"
# (4) full capture of fb_em_kernel on the single-GPU-sized cut of the same workload (c4s: same reads / gap lengths, 500 gaps):
#     a --set full pass replays every launch ~40 times, which the multi-second launches of c4 do not allow inside the GPU budget
python bench.py --workload c4s --gpus 1 --steps 1 --warmup 1 --no-cpu-baseline > $O/r2_bench_c4s_n1.json 2> $O/r2_bench_c4s_n1.err || tail -5 $O/r2_bench_c4s_n1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fb_em_kernel --launch-skip 4 -c 10 -f -o $O/r2_prof_c4s python bench.py --workload c4s --gpus 1 --steps 1 --warmup 0 --no-cpu-baseline > $O/r2_ncu_full.log 2>&1
echo "full capture rc=$?"
ncu -i $O/r2_prof_c4s.ncu-rep --page raw --csv > $O/r2_fb_em_kernel_c4s_raw.csv 2>/dev/null
ls -la $O | tail -20
