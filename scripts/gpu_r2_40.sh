#!/bin/bash
# Round 2, GPU call 40 (1 GPU): --set full capture (with source) of the stabilised tracer / momentum element kernels after the hoist.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_advdiff_stage_kernel -s 1 -c 1 -o gpurun_out/r2_su2_adv_stage -f python scripts/prof_su.py 64 > gpurun_out/r2_40_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_momentum_stage_kernel -s 1 -c 1 -o gpurun_out/r2_su2_mom_stage -f python scripts/prof_su.py 64 > gpurun_out/r2_40_m.log 2>&1
ls -la gpurun_out/r2_su2_*.ncu-rep
