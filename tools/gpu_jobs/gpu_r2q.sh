#!/bin/bash
# round 2, late: full GPU suite, racecheck of the final resident kernel + the LU panel exchange + the ring stencils, full bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -8 gpurun_out/r2q_pytest.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanity_resident.py 100 7 > gpurun_out/r2q_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2q_racecheck.txt
tail -4 gpurun_out/r2q_racecheck.txt
python bench.py --steps 2 --warmup 1 > gpurun_out/r2q_bench_full.json 2> gpurun_out/r2q_bench_full.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2q_bench_full.json; tail -3 gpurun_out/r2q_bench_full.err
