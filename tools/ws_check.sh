#!/bin/bash
# Parity of the persistent warp-specialised GEMMs, one kernel at a time (a deadlock only costs its own timeout).
out=gpurun_out/${1:-ws}
mkdir -p $out
for k in project expand dc dx all; do
  TFNAS_WS=$k timeout 300 python -m pytest tests/test_mixedop_gpu.py -x -q > $out/pytest_$k.log 2>&1
  echo "TFNAS_WS=$k exit $?"; tail -3 $out/pytest_$k.log
done
