#!/bin/bash
# GPU box: what the driver runs at round end -- gpu tests, smoke(), the default bench and the reference arm.  usage: tools/gpu_final.sh <tag>
tag=${1:-final}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8 | tee gpurun_out/pytest_$tag.log
( time python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) 2>&1 | tail -5 | tee gpurun_out/smoke_$tag.log
( time python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err ) 2>&1 | tail -4
head -c 600 gpurun_out/bench_$tag.json; echo
( time python bench.py --impl reference > gpurun_out/benchref_$tag.json 2> gpurun_out/benchref_$tag.err ) 2>&1 | tail -4
head -c 900 gpurun_out/benchref_$tag.json; echo
