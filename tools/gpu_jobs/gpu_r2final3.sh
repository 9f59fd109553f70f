#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 > gpurun_out/r2final_bench_full.json 2> gpurun_out/r2final_bench_full.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/r2final_bench_full.json
