#!/bin/bash
# round 2, second GPU call: the whole GPU suite (new features), the full bench line with every leg, memcheck of the new kernels
mkdir -p gpurun_out
SKIP=""
[ -f tests/golden/bruss3d_n100_newton.npz ] || SKIP='-k not(config3_full_solve)'
python -m pytest tests -m gpu -q $SKIP > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
python bench.py --steps 2 --warmup 1 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?" >> gpurun_out/r2b_bench.err
timeout 600 compute-sanitizer --tool memcheck python tools/sanity_round2.py > gpurun_out/r2b_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2b_memcheck.txt
tail -25 gpurun_out/r2b_pytest.log; tail -3 gpurun_out/r2b_bench.err; cut -c1-300 gpurun_out/r2b_bench.json; tail -4 gpurun_out/r2b_memcheck.txt
