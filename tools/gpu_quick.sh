#!/bin/bash
# GPU box: engine parity tests + product tests + a short bench (no ncu).  usage: tools/gpu_quick.sh <tag>
tag=${1:-q}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_$tag.log
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || { echo "bench failed"; tail -20 gpurun_out/bench_$tag.err; exit 1; }
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("value %.3e e2e %.3e ms/step %.0f launches %d kernel_ms/launch %.2f frac %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"]))
print(d["e2e"]["host_seconds_per_step"]); print({k: d["roofline"].get(k) for k in ("smem", "traffic")})
PY
