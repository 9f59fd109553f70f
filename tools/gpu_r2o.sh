#!/bin/bash
mkdir -p gpurun_out
B200_LU_TRACE=1 python tools/lu_bench.py 128 > gpurun_out/r2o_lu.txt 2>&1
grep "lu trace" gpurun_out/r2o_lu.txt | tail -12
