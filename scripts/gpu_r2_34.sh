#!/bin/bash
# Round 2, GPU call 34 (1 GPU): first ncu look at the 2-D kernels (driven_cavity / lock_exchange option sets, 2048^2 x 2 triangles).
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,sm__warps_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:staged_ -c 6 --csv --log-file gpurun_out/r2_launches_2d.csv python scripts/ab_kernels.py 2048 2d > gpurun_out/r2_34_2d.log 2>&1
for k in momentum advdiff; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:staged_${k}_kernel -s 2 -c 1 -o gpurun_out/r2_2d_$k -f python scripts/ab_kernels.py 2048 2d > gpurun_out/r2_34_ncu_$k.log 2>&1
done
ls -la gpurun_out/r2_2d_*.ncu-rep
