#!/bin/bash
# Round 2, GPU call 5 (1 GPU): parity (fused call, per-component absorption pass, micro-optimised kernels), bench N=1 both step flavours, launch metrics.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_5_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_5_pytest.log
tail -4 gpurun_out/r2_5_pytest.log
timeout 1200 python bench.py > gpurun_out/r2_5_bench_n1.json 2> gpurun_out/r2_5_bench_n1.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_5_bench_n1.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "sep", d["separate_kernels_ms_rank0"], "fused", d["fused_kernel_ms_rank0"], "frac", d["roofline"]["frac"], d["roofline"]["tracer"]["frac"], "setup", d["setup_s"])
for c in d["configs"]: print(c["config"][:44], round(c["momentum_ms"],3), round(c["tracer_ms"],3), round(c["gel_s"],2), c["momentum_path"], c["tracer_path"])
PY
timeout 600 python bench.py --step fused --no-cpu-baseline --no-configs --no-e2e > gpurun_out/r2_5_bench_n1_fused.json 2> gpurun_out/r2_5_bench_n1_fused.err; cut -c1-200 gpurun_out/r2_5_bench_n1_fused.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio \
  --clock-control none -k regex:staged_ -c 9 --csv --log-file gpurun_out/r2_5_launches_128.csv python bench.py --cells 128 --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-e2e > gpurun_out/r2_5_ncu.log 2>&1
tail -2 gpurun_out/r2_5_ncu.log | cut -c1-200
