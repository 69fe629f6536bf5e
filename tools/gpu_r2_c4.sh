#!/bin/bash
# GPU box: the C4 bench at N GPUs (N = $1, default 1).  usage: tools/gpu_r2_c4.sh <N> <tag> [steps] [warmup]
N=${1:-1}; tag=${2:-c4}; steps=${3:-2}; warm=${4:-1}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nproc; free -g | head -2
if [ "$N" = 1 ]; then
  python bench.py --gpus 1 --steps $steps --warmup $warm --no-cpu-baseline > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err || { echo "bench failed"; tail -30 gpurun_out/bench_${tag}_n1.err; exit 1; }
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $steps --warmup $warm > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err || { echo "bench failed"; tail -30 gpurun_out/bench_${tag}_n$N.err; exit 1; }
fi
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${tag}_n$N.json"))
print("N=$N value %.3e e2e %.3e s/step %.2f gaps/s %.0f frac %.4f prep %.0fs launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"]/1e3, d["e2e"]["gaps_per_s"], d["roofline"]["frac"], d["data_prep_s"], d["gpu_launches"]))
print(d["e2e"]["seconds_per_step_breakdown"]); print(d["e2e"]["host_seconds_per_step"]); print(d["e2e"]["device_ms_per_gpu_per_step"]); print(d["clocks"]); print("d2h MB/step", d["e2e"]["d2h_bytes_per_step"]/1e6)
PY
