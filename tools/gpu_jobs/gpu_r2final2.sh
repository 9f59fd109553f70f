#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2final_pytest.log
tail -4 gpurun_out/r2final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --steps 2 --warmup 3 > gpurun_out/r2final_bench_full.json 2> gpurun_out/r2final_bench_full.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/r2final_bench_full.json
