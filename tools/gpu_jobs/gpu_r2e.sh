#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2e_gpus.txt
./tests/abi_c/abi_c_check > gpurun_out/r2e_abi_c.log 2>&1; echo "abi_c rc=$?" >> gpurun_out/r2e_abi_c.log
python -m pytest tests -m gpu -q -k "two_devices or abi_c" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 1 --warmup 1 --no-legs > gpurun_out/r2e_bench_2gpu.json 2> gpurun_out/r2e_bench_2gpu.err; echo "bench rc=$?" >> gpurun_out/r2e_bench_2gpu.err
cat gpurun_out/r2e_abi_c.log; tail -3 gpurun_out/r2e_pytest.log; tail -5 gpurun_out/r2e_bench_2gpu.err; cut -c1-200 gpurun_out/r2e_bench_2gpu.json
