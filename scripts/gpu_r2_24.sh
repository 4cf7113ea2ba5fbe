#!/bin/bash
# Round 2, GPU call 24 (1 GPU): where the SU two-pass path spends its time (launch list + full capture of the element kernels).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,launch__registers_per_thread --clock-control none -s 6 -c 8 --csv --log-file gpurun_out/r2_24_su_launches.csv python scripts/prof_su.py 64 2>&1 | tail -1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gather_advdiff_stage -s 1 -c 1 -o gpurun_out/r2_su_adv_stage -f python scripts/prof_su.py 64 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gather_rows_kernel -s 2 -c 1 -o gpurun_out/r2_su_rows -f python scripts/prof_su.py 64 > /dev/null 2>&1
ls -la gpurun_out/r2_su_*
