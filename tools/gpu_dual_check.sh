#!/bin/bash
# GPU box: run the host program on the dual engine (CUDA + oracle side by side) over a golden fixture and the c2 bench case
cd "$(dirname "$0")/.."
export FB_DUAL_DEVICE_LIB=$PWD/figbird_b200/_build/libfigbird_b200.so FB_DUAL_ORACLE_LIB=$PWD/oracle/_build/libfb_oracle.so FIGBIRD_QUIET=1
rm -rf /tmp/dual && mkdir -p /tmp/dual && tar -xzf tests/golden/g2.tar.gz -C /tmp/dual
for mode in partial unmapped; do
  c=/tmp/dual/g2; mkdir -p $c/run_$mode/Temp; cp $c/$mode/Temp/{gapInfo,stat,stat2}.txt $c/run_$mode/Temp/
  if [ $mode = partial ]; then a="200 100 1 1 0"; else a="230 100 2 0 1"; fi
  echo "== g2 $mode"; oracle/_build/fillgaps_dual $c/draft.fa $a 2 $c/$mode/myout.sam $c/run_$mode/Temp/ $c/$mode/Gaps/ 30 100 0 0 200 2>&1 | tail -25
done
