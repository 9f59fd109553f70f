#!/bin/bash
# round 2, third GPU call: full GPU suite, racecheck + memcheck of the final binary, ncu capture of the final resident kernel
mkdir -p gpurun_out
SKIP=""
[ -f tests/golden/bruss3d_n100_newton.npz ] || SKIP='-k not(config3_full_solve)'
python -m pytest tests -m gpu -q $SKIP > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
python tools/resident_phases.py 100 300 mgs resident > gpurun_out/r2c_phases.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanity_resident.py 100 5 > gpurun_out/r2c_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2c_racecheck.txt
timeout 600 compute-sanitizer --tool memcheck python tools/sanity_round2.py > gpurun_out/r2c_memcheck_round2.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2c_memcheck_round2.txt
timeout 600 compute-sanitizer --tool memcheck python tools/sanity_small.py > gpurun_out/r2c_memcheck_small.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2c_memcheck_small.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resident3g_arnoldi_kernel -s 1400 -c 1 -o gpurun_out/r2c_r3g_full python tools/explore_resident.py 100 resident > gpurun_out/r2c_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 120 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 1 --warmup 1 --no-ensemble --no-legs > gpurun_out/r2c_launch_bench.log 2>&1
tail -6 gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_phases.log; tail -3 gpurun_out/r2c_racecheck.txt; tail -3 gpurun_out/r2c_memcheck_round2.txt; tail -3 gpurun_out/r2c_memcheck_small.txt
