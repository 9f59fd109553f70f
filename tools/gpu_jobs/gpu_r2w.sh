#!/bin/bash
# the driver's SCALE command at N = $1
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/r2w_gpus_$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r2w_bench_${N}gpu.json 2> gpurun_out/r2w_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2w_bench_${N}gpu.json')); e=d['ensemble']
print('n_gpus',d['n_gpus'],'value',d['value'],'ms_per_step',d['ms_per_step'],'e2e',d['e2e']['value'])
print('ensemble',e['problems_per_s'],e['ms'],e.get('solve_ms_per_rank'),e.get('tail_imbalance'),e.get('collective_ms'),e.get('collectives'))
PY
tail -3 gpurun_out/r2w_bench_${N}gpu.err
