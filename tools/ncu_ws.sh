#!/bin/bash
# ncu --set full (with source counters) on the persistent GEMM kernels of ONE alpha-mode MixedOP fwd+bwd.
# usage: bash tools/ncu_ws.sh <tag> <block index> [kernel regex]
tag=${1:-ws}; blk=${2:-1}; rx=${3:-k_ws_}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:$rx" --launch-skip 4 --launch-count 4 \
  -o $out/ws_blk${blk} -f python tools/time_mixedops.py --mode alpha --only $blk --reps 1 --out $out/ncu_dummy.json > $out/ncu_blk${blk}.log 2>&1
echo "ncu exit $?"
ncu -i $out/ws_blk${blk}.ncu-rep --page raw --csv > $out/ws_blk${blk}_raw.csv 2>> $out/ncu_blk${blk}.log
ls -la $out
