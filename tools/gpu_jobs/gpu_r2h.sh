#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/sanity_round2.py > gpurun_out/r2h_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2h_memcheck.log
tail -12 gpurun_out/r2h_memcheck.log
python tools/stencil_bench.py 100 80 > gpurun_out/r2h_stencil.json 2> gpurun_out/r2h_stencil.err
cat gpurun_out/r2h_stencil.json | tr -d '\n ' ; echo
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:bruss3d -c 90 --csv --log-file gpurun_out/r2h_stencil_ncu.csv python tools/stencil_bench.py 100 > /dev/null 2>&1
python tools/mgs_attainable.py 100 > gpurun_out/r2h_mgs_attainable.json 2>&1; cat gpurun_out/r2h_mgs_attainable.json | tr -d '\n '; echo
python -m pytest tests -m gpu -q --deselect tests/test_gpu_n100_parity.py::test_config3_full_solve_vs_cpu_port_artefact > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -15 gpurun_out/r2h_pytest.log
python bench.py --steps 2 --warmup 1 --no-legs > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; cut -c1-1500 gpurun_out/r2h_bench.json
