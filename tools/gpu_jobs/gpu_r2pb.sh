#!/bin/bash
mkdir -p gpurun_out
python tools/precond_breakdown.py > gpurun_out/r2pb.json 2>&1; cat gpurun_out/r2pb.json | tr -d '\n ' | cut -c1-1500; echo
