#!/bin/bash
mkdir -p gpurun_out
python tools/stencil_bench.py 100 80 > gpurun_out/r2g_stencil.json 2> gpurun_out/r2g_stencil.err
cat gpurun_out/r2g_stencil.json | tr -d '\n ' ; echo
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:bruss3d -c 60 --csv --log-file gpurun_out/r2g_stencil_ncu.csv python tools/stencil_bench.py 100 > /dev/null 2>&1
python -m pytest tests -m gpu -q --deselect tests/test_gpu_n100_parity.py::test_config3_full_solve_vs_cpu_port_artefact > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -25 gpurun_out/r2g_pytest.log
