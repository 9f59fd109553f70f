#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "lu or dense or qr or levenberg or broyden or abi_c or newton_configs or julia" > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log
tail -5 gpurun_out/r2s_pytest.log
B200_LU_TRACE=1 timeout 300 python tools/lu_bench.py 128 > gpurun_out/r2s_lu.txt 2>&1
grep "lu trace" gpurun_out/r2s_lu.txt | tail -10; tail -1 gpurun_out/r2s_lu.txt | cut -c1-200
timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python tools/sanity_stencil.py > gpurun_out/r2r_racecheck_stencil.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2r_racecheck_stencil.txt
tail -5 gpurun_out/r2r_racecheck_stencil.txt
