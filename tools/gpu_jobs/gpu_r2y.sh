#!/bin/bash
# final binary: racecheck of the resident kernel (small grids + a burst of 4 full-size steps), full bench, ncu re-capture
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanity_resident.py 100 4 > gpurun_out/r2y_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2y_racecheck.txt
tail -4 gpurun_out/r2y_racecheck.txt
python bench.py --steps 2 --warmup 3 > gpurun_out/r2y_bench_full.json 2> gpurun_out/r2y_bench_full.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2y_bench_full.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 120 --csv --log-file gpurun_out/r2y_launches_bench_N100.csv python bench.py --steps 1 --warmup 1 --no-legs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:resident3g_arnoldi_kernel -s 1400 -c 1 -o gpurun_out/r2y_resident3g python tools/explore_resident.py 100 resident > gpurun_out/r2y_explore.log 2>&1
ls -la gpurun_out/r2y_*
