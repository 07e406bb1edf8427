#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, per-shape kernel timings, one full ncu capture.
# usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag] [stages]
#   stages: any of t(ests) b(ench) l(aunch list) k(ernel timings) n(cu full) d(ram traffic per kernel), default "tblkn"
tag=${1:-r1}
stages=${2:-tblkn}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/gpu.txt 2>&1
nproc > $out/nproc.txt
if [[ $stages == *t* ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $out/pytest_gpu.log
  tail -5 $out/pytest_gpu.log
fi
if [[ $stages == *b* ]]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err
  echo "bench exit $?"; cat $out/bench.json | cut -c1-600
fi
if [[ $stages == *k* ]]; then
  timeout 600 python tools/time_mixedops.py --mode alpha --out $out/kernels_alpha.json > $out/kernels_alpha.txt 2>&1
  timeout 600 python tools/time_mixedops.py --mode single --out $out/kernels_single.json > $out/kernels_single.txt 2>&1
  tail -30 $out/kernels_alpha.txt
fi
if [[ $stages == *l* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches.csv \
    python bench.py --profile-only --steps 1 --warmup ${NCU_WARMUP:-1} --no-cpu-baseline > $out/launches.log 2>&1
  echo "ncu launch list exit $?"
  python tools/summarize_ncu.py $out/launches.csv > $out/launches_summary.csv 2>> $out/launches.log
  head -30 $out/launches_summary.csv
fi
if [[ $stages == *n* ]]; then
  # full-set capture of the GEMM + depthwise kernels of ONE alpha-mode MixedOP fwd+bwd (second pass; the first is warm-up)
  for blk in ${NCU_BLOCKS:-1 10}; do
    timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_ws_|k_um_(expand|project|dc|dx)|k_dw' \
      --launch-skip 8 --launch-count 8 -o $out/mixedop_blk${blk} -f \
      python tools/time_mixedops.py --mode alpha --only $blk --reps 1 --out $out/ncu_dummy.json > $out/ncu_full_blk${blk}.log 2>&1
    echo "ncu full blk $blk exit $?"
    ncu -i $out/mixedop_blk${blk}.ncu-rep --page raw --csv > $out/mixedop_blk${blk}_raw.csv 2>> $out/ncu_full_blk${blk}.log
    ncu -i $out/mixedop_blk${blk}.ncu-rep --page details --csv > $out/mixedop_blk${blk}_details.csv 2>> $out/ncu_full_blk${blk}.log
    sz=$(stat -c %s $out/mixedop_blk${blk}.ncu-rep 2>/dev/null || echo 0)
    if [ "$sz" -gt 12000000 ]; then rm -f $out/mixedop_blk${blk}.ncu-rep; echo "dropped ncu-rep ($sz bytes)"; fi
  done
fi
if [[ $stages == *d* ]]; then
  # DRAM bytes per launch of the GEMM / depthwise kernels over one search unit (two metrics: a single ncu pass per kernel)
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $out/dram.csv \
    -k 'regex:k_ws_|k_um_|k_dw|k_dxfin|k_b2b' python bench.py --profile-only --steps 1 --warmup ${NCU_WARMUP:-1} --no-cpu-baseline > $out/dram.log 2>&1
  echo "ncu dram exit $?"
  python tools/ncu_traffic.py $out/dram.csv > $out/ncu_traffic.json 2>> $out/dram.log
  head -c 600 $out/ncu_traffic.json
fi
du -sh gpurun_out; ls -la $out
