#!/bin/bash
mkdir -p gpurun_out
B200_LU_TRACE=1 timeout 300 python tools/lu_bench.py 128 > gpurun_out/r2t_lu.txt 2>&1
grep "lu trace" gpurun_out/r2t_lu.txt | tail -10; tail -1 gpurun_out/r2t_lu.txt | cut -c1-200
timeout 300 python -m pytest tests -m gpu -q -k "lu or dense or broyden or limited" 2>&1 | tail -3
