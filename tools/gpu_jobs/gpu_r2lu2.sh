#!/bin/bash
# after the grouped tile order / fused U12 / LL panel: DMMA %, DRAM bytes and L2 hit rate of the first big trailing updates
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,lts__t_bytes.sum --clock-control none -k regex:gemm_sub_w8 -c 40 --csv --log-file gpurun_out/r2lu2_gemm_launches.csv python tools/lu_bench.py 128 > gpurun_out/r2lu2.log 2>&1
tail -1 gpurun_out/r2lu2.log | cut -c1-200
