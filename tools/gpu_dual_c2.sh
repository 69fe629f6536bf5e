#!/bin/bash
# GPU box: the c2 bench case through the dual engine (CUDA vs oracle item by item); needs oracle/_ref for input prep
cd "$(dirname "$0")/.."
export FB_DUAL_DEVICE_LIB=$PWD/figbird_b200/_build/libfigbird_b200.so FB_DUAL_ORACLE_LIB=$PWD/oracle/_build/libfb_oracle.so FIGBIRD_QUIET=1
python - <<'PY'
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import bench
bench.prepare_case('/tmp/fb_bench/c2_rank0', bench.WORKLOADS['c2'], 102)
PY
c=/tmp/fb_bench/c2_rank0
for mode in ${1:-partial unmapped}; do
  rm -rf $c/dual_$mode; mkdir -p $c/dual_$mode/Temp; cp $c/$mode/Temp/{gapInfo,stat,stat2}.txt $c/dual_$mode/Temp/
  if [ $mode = partial ]; then a="200 100 1 1 0"; else a="230 100 2 0 1"; fi
  echo "== c2 $mode"; ( time oracle/_build/fillgaps_dual $c/draft.fa $a 16 $c/$mode/myout.sam $c/dual_$mode/Temp/ $c/$mode/Gaps/ 30 100 0 0 200 ) 2>&1 | tail -30
  cp $c/dual_$mode/Temp/gapout.txt gpurun_out/dual_gapout_$mode.txt
done
