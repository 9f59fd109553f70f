#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,lts__t_bytes.sum --clock-control none -k regex:gemm_sub_w8 -c 260 --csv --log-file gpurun_out/r2m_lu_gemm_launches.csv python tools/lu_bench.py 128 > gpurun_out/r2m_lu.log 2>&1
tail -2 gpurun_out/r2m_lu.log
