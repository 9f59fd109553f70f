#!/bin/bash
# round 2, final single-GPU pass: full GPU suite, smoke(), full bench line, launch list + one --set full capture of the resident kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
tail -5 gpurun_out/r2v_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 2 --warmup 3 > gpurun_out/r2v_bench_full.json 2> gpurun_out/r2v_bench_full.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2v_bench_full.json; tail -2 gpurun_out/r2v_bench_full.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2v_bench_reference.json 2> gpurun_out/r2v_bench_reference.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r2v_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 120 --csv --log-file gpurun_out/r2v_launches_bench_N100.csv python bench.py --steps 1 --warmup 1 --no-legs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:resident3g_arnoldi_kernel -s 1400 -c 1 -o gpurun_out/r2v_resident3g python tools/explore_resident.py 100 resident > gpurun_out/r2v_explore.log 2>&1
ls -la gpurun_out/r2v_*
