#!/bin/bash
mkdir -p gpurun_out
python tools/stencil_bench.py 100 80 > gpurun_out/r2j_stencil.json 2> gpurun_out/r2j_stencil.err
cat gpurun_out/r2j_stencil.json | tr -d '\n ' ; echo
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"bruss3d|copy" -c 120 --csv --log-file gpurun_out/r2j_stencil_ncu.csv python tools/stencil_bench.py 100 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_sub_w8 -s 40 -c 1 -o gpurun_out/r2j_lu_gemm python tools/lu_bench.py 128 > gpurun_out/r2j_lu.log 2>&1
tail -3 gpurun_out/r2j_lu.log
