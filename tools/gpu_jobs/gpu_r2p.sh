#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "lu or dense or qr or levenberg or broyden or abi_c" > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -6 gpurun_out/r2p_pytest.log
B200_LU_TRACE=1 timeout 300 python tools/lu_bench.py 128 > gpurun_out/r2p_lu.txt 2>&1
grep "lu trace" gpurun_out/r2p_lu.txt | tail -11; tail -1 gpurun_out/r2p_lu.txt | cut -c1-300
