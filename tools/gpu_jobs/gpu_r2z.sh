#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "ensemble or abi_c or smoke" 2>&1 | tail -3
python bench.py --steps 1 --warmup 1 --no-legs > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2z_bench.json')); e=d['ensemble']
print('headline',d['value'],d['roofline']['frac'],'ensemble',e['problems_per_s'],e['ms'],e['n_success'],e['worst_resid_inf'],e['gmres_jvps'])
PY
