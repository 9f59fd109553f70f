#!/bin/bash
mkdir -p gpurun_out
for p in 148 112 80 64; do
  echo "PMAX $p" >> gpurun_out/r2n_lu.txt
  B200_LU_PMAX=$p python tools/lu_bench.py 128 2>&1 | tail -1 | cut -c1-400 >> gpurun_out/r2n_lu.txt
done
cat gpurun_out/r2n_lu.txt
