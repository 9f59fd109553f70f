#!/bin/bash
mkdir -p gpurun_out
for v in 4 5; do B200_LU_GEMM2=$v python tools/lu_bench.py 128 > gpurun_out/r2d_lu_v$v.log 2>&1; done
B200_LU_NBO=768 python tools/lu_bench.py 128 > gpurun_out/r2d_lu_nbo768.log 2>&1
python -m pytest tests -m gpu -q -k "lu or dense or getrf" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
cat gpurun_out/r2d_lu_v4.log gpurun_out/r2d_lu_v5.log gpurun_out/r2d_lu_nbo768.log; tail -3 gpurun_out/r2d_pytest.log
