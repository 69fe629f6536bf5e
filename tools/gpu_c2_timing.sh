#!/bin/bash
# GPU box: per-phase timings of the product fillgaps on the c2 bench case, for a few host configurations
cd "$(dirname "$0")/.."
export FIGBIRD_QUIET=1
python - <<'PY'
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import bench
bench.prepare_case('/tmp/fb_bench/c2_rank0', bench.WORKLOADS['c2'], 102)
PY
c=/tmp/fb_bench/c2_rank0
run() {  # mode label env...
  mode=$1; shift; label=$1; shift
  rm -rf $c/t_$mode; mkdir -p $c/t_$mode/Temp; cp $c/$mode/Temp/{gapInfo,stat,stat2}.txt $c/t_$mode/Temp/
  if [ $mode = partial ]; then a="200 100 1 1 0"; else a="230 100 2 0 1"; fi
  s=$(date +%s.%N)
  env "$@" FIGBIRD_METRICS=$c/t_$mode/m.json figbird_b200/_build/fillgaps $c/draft.fa $a 16 $c/$mode/myout.sam $c/t_$mode/Temp/ $c/$mode/Gaps/ 30 100 0 0 200 > /dev/null
  e=$(date +%s.%N)
  python -c "
import json; m=json.load(open('$c/t_$mode/m.json'))
print('%-9s %-22s wall %.2f | load %.2f model %.2f prep %.2f fill %.2f (ctx %.2f workers %.2f cpu %.2f | engine calls %.2f, copy %.2f, device %.3f) write %.2f | launches %d ticks %d d2h %.0f MB' % ('$mode','$label',$e-$s,m['t_load'],m['t_model'],m['t_prepare'],m['t_fill'],m['t_ctx_upload'],m['t_workers'],m['cpu_workers'],m['t_engine_calls'],m['t_result_copy'],m['device_ms']/1e3,m['t_write'],m['kernel_launches'],m['ticks'],m['d2h_bytes']/1e6))"
}
for mode in partial unmapped; do
  run $mode warmup X=1
  run $mode default X=1
  run $mode default2 X=1
done
