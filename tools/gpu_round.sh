#!/bin/bash
# GPU box: gpu tests, bench, ncu launch list and one full capture of the dominant kernel.  usage: tools/gpu_round.sh <tag> [skip_tests]
tag=${1:-rX}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ -z "$2" ]; then
  python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_$tag.log
fi
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || { echo "bench failed"; tail -20 gpurun_out/bench_$tag.err; exit 1; }
cat gpurun_out/bench_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_$tag.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fb_em_kernel --launch-skip 6 -c 4 -f -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -12
