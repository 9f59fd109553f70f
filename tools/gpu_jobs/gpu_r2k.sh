#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "residual_jvp_vjp or stencil_kernels or jvp_properties or z_invariance" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -5 gpurun_out/r2k_pytest.log
python tools/stencil_bench.py 100 80 > gpurun_out/r2k_stencil.json 2> gpurun_out/r2k_stencil.err
cat gpurun_out/r2k_stencil.json | tr -d '\n ' ; echo
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__cycles_active.avg --clock-control none -k regex:"bruss3d|copy|axpby" -c 160 --csv --log-file gpurun_out/r2k_stencil_ncu.csv python tools/stencil_bench.py 100 > /dev/null 2>&1
