#!/bin/bash
# Round 2, GPU call 17 (1 GPU): kmk parity, full suite again, racecheck of the GATHER surface test, shared-memory bank
# conflicts of the staged momentum kernel on a box / renumbered / Delaunay mesh.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_17_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_17_pytest.log
tail -6 gpurun_out/r2_17_pytest.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_surface_gpu.py -m gpu -q -k "cube-parallel and gather" > gpurun_out/r2_17_racecheck.log 2>&1
grep -E "hazard|ERROR SUMMARY|passed|failed" gpurun_out/r2_17_racecheck.log | head -8
M=l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio
for cfg in "box 64" "shuffled 64" "delaunay 250000"; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:staged_momentum -c 1 --csv --log-file gpurun_out/r2_17_ncu_$(echo $cfg | tr ' ' '_').csv python scripts/prof_unstructured.py $cfg 2>&1 | tail -1
done
