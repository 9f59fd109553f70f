#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python tools/sanity_stencil.py > gpurun_out/r2r_racecheck_stencil.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2r_racecheck_stencil.txt
tail -6 gpurun_out/r2r_racecheck_stencil.txt
python tools/stencil_bench.py 100 > gpurun_out/r2r_stencil.json 2>/dev/null; cat gpurun_out/r2r_stencil.json | tr -d '\n ' | cut -c1-400; echo
