#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "precond or multigrid or mg or precs" 2>&1 | tail -3
python tools/precond_breakdown.py > gpurun_out/r2pc.json 2>&1; cat gpurun_out/r2pc.json | tr -d '\n ' | cut -c1-900; echo
timeout 300 compute-sanitizer --tool memcheck python tools/sanity_round2.py 2>&1 | tail -3
