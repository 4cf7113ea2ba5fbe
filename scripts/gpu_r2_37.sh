#!/bin/bash
# Round 2, GPU call 37 (1 GPU): final captures with the searched strips and the occupancy classes -- full GPU suite, smoke,
# S3 launch list with DRAM bytes (traffic), launch list at 128^3, --set full captures (96^3) of the momentum / tracer / fused /
# absorption kernels, the default bench line and the reference arm.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_37_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_37_pytest.log; tail -3 gpurun_out/r2_37_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_37_smoke.log 2>&1; tail -2 gpurun_out/r2_37_smoke.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none -k regex:staged_ -c 12 --csv --log-file gpurun_out/r2_launches_S3_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-e2e > gpurun_out/r2_37_ncu_S3.log 2>&1; tail -1 gpurun_out/r2_37_ncu_S3.log | cut -c1-160
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_128_final.csv python bench.py --cells 128 --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-e2e > gpurun_out/r2_37_ncu_128.log 2>&1; tail -1 gpurun_out/r2_37_ncu_128.log | cut -c1-160
for k in momentum advdiff fused; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:staged_${k}_kernel -s 2 -c 1 -o gpurun_out/r2_final_$k -f python bench.py --cells 96 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/r2_37_ncu_$k.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:staged_momentum_absorb -s 1 -c 1 -o gpurun_out/r2_final_absorb -f python scripts/prof_absorb.py 96 > gpurun_out/r2_37_ncu_absorb.log 2>&1
ls -la gpurun_out/r2_final_*.ncu-rep
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_37_ref_n1.json 2> gpurun_out/r2_37_ref_n1.err; echo "ref exit $?"; cut -c1-300 gpurun_out/r2_37_ref_n1.json
timeout 1500 python bench.py > gpurun_out/r2_37_bench_n1.json 2> gpurun_out/r2_37_bench_n1.err; echo "bench exit $? wall ${SECONDS}s"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_37_bench_n1.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "sep", d["separate_kernels_ms_rank0"], "fused", d["fused_kernel_ms_rank0"], "frac", d["roofline"]["frac"], d["roofline"]["tracer"]["frac"], "mom", d["roofline"]["kernel_ms"], "tracer", d["roofline"]["tracer"]["kernel_ms"], "e2e", d["e2e"]["value"], "setup", d["setup_s"])
for c in d["configs"]:
    if "error" in c: print(c); continue
    print(c["config"][:60], round(c["momentum_ms"],3), round(c["tracer_ms"],3), round(c["gel_s"],2), {k: round(v,3) for k,v in c.get("roofline_frac",{}).items() if k!="algorithmic_bytes_per_element"})
PY
