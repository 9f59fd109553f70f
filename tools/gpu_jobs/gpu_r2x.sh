#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_resident.py tests/test_gpu_n100_parity.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -4
python bench.py --steps 2 --warmup 1 --no-legs > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2x_bench.json'))
print('headline',d['value'],d['roofline']['frac'],d['ms_per_step'],'e2e',d['e2e']['value'],'ms/launch',d['roofline']['ms_per_launch'],d['resid_inf'],d['config']['step'])
PY
