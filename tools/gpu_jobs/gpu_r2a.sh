#!/bin/bash
# round 2, first GPU call: parity suite, A/B of the resident kernels, racecheck of the new default, ncu capture
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "not config3_full_solve" > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
for v in 1 2; do
  B200_RS3=$v python tools/resident_phases.py 100 300 mgs resident > gpurun_out/r2a_phases_rs3_$v.log 2>&1
  B200_RS3=$v python bench.py --steps 2 --warmup 1 --no-ensemble > gpurun_out/r2a_bench_rs3_$v.json 2> gpurun_out/r2a_bench_rs3_$v.err
done
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanity_resident.py 100 7 > gpurun_out/r2a_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2a_racecheck.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resident3g_arnoldi_kernel -s 1400 -c 1 -o gpurun_out/r2a_r3g_full python tools/explore_resident.py 100 resident > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_bench_rs3_*.json | cut -c1-400; tail -5 gpurun_out/r2a_racecheck.txt
