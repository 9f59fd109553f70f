#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/r2pd.json 2>gpurun_out/r2pd.err
import json, sys
sys.path.insert(0, ".")
import torch, bench
import nonlinearsolve_jl_b200 as nls
ctx = nls.Context(0)
r = bench.leg_precond(nls, torch, ctx)
print(json.dumps(r))
PY
cut -c1-1200 gpurun_out/r2pd.json; tail -3 gpurun_out/r2pd.err
