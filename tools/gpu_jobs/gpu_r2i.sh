#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "broyden or gmres_vs_oracle or radius_schemes or vector_ops" > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -12 gpurun_out/r2i_pytest.log
ncu --set full --clock-control none --import-source on -k regex:bruss3d_ring -s 30 -c 2 -o gpurun_out/r2i_ring python tools/stencil_bench.py 100 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
