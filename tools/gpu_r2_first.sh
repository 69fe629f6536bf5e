#!/bin/bash
# GPU box, round 2 first call: parity tests, then kernel build variants A/B on c2, then c4s with the product build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee gpurun_out/r2a_smi.txt; nproc | tee -a gpurun_out/r2a_smi.txt; free -g | head -2 | tee -a gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_r2a.log
bash tools/gpu_variants.sh base c3 t256c4 t256c3 fuse1 fuse12 wtr fin2 2>&1 | tee gpurun_out/variants_r2a.log
python bench.py --workload c4s --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2a_c4s.json 2> gpurun_out/bench_r2a_c4s.err; tail -c 1500 gpurun_out/bench_r2a_c4s.json
