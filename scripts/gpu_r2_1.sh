#!/bin/bash
# Round 2, GPU call 1: parity of the rewritten staged STRIP kernels (byte-offset plan entries, compile-time chunk stride,
# tracer absorption/source), A/B against the round-1 library on the same box, set-up time, launch metrics.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_1_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_1_pytest.log
tail -5 gpurun_out/r2_1_pytest.log
CGASM_DEBUG=1 timeout 600 python scripts/sweep_strip.py --cells 256 --reps 10 --configs default= --tag _new > gpurun_out/r2_1_sweep_new.log 2>&1; tail -4 gpurun_out/r2_1_sweep_new.log
CGASM_LIB=$PWD/ab/libcgasm_r1.so timeout 900 python scripts/sweep_strip.py --cells 256 --reps 10 --configs default= --tag _r1 > gpurun_out/r2_1_sweep_r1.log 2>&1; tail -3 gpurun_out/r2_1_sweep_r1.log
timeout 600 python scripts/sweep_strip.py --cells 128 --reps 10 --configs default= --shuffle --tag _shuf > gpurun_out/r2_1_sweep_shuf.log 2>&1; tail -3 gpurun_out/r2_1_sweep_shuf.log
timeout 600 python scripts/sweep_strip.py --cells 128 --reps 10 --configs default= --tag _box > gpurun_out/r2_1_sweep_box128.log 2>&1; tail -2 gpurun_out/r2_1_sweep_box128.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:staged_ -c 8 --csv --log-file gpurun_out/r2_1_launches_128.csv python scripts/sweep_strip.py --cells 128 --reps 2 --configs default= --tag _ncu > gpurun_out/r2_1_ncu.log 2>&1
tail -9 gpurun_out/r2_1_launches_128.csv | cut -c1-400
