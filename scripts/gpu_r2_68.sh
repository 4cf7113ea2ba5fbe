#!/bin/bash
# Round 2, GPU call 68 (1 GPU): final library (diagonal by row sum, landed uploads): full GPU suite, smoke, default bench line.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_68_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_68_pytest.log; tail -3 gpurun_out/r2_68_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_68_smoke.log 2>&1; tail -2 gpurun_out/r2_68_smoke.log
timeout 1500 python bench.py > gpurun_out/r2_68_bench_n1.json 2> gpurun_out/r2_68_bench_n1.err; echo "bench exit $? wall ${SECONDS}s"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_68_bench_n1.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "sep", d["separate_kernels_ms_rank0"], "fused", d["fused_kernel_ms_rank0"], "frac", d["roofline"]["frac"], d["roofline"]["tracer"]["frac"], d["roofline"]["combined_frac"], "mom", d["roofline"]["kernel_ms"], "tracer", d["roofline"]["tracer"]["kernel_ms"], "e2e", d["e2e"]["value"], "setup", d["setup_s"])
for c in d["configs"]:
    if "error" in c: print(c); continue
    print(c["config"][:60], round(c["momentum_ms"],3), round(c["tracer_ms"],3), round(c["gel_s"],2), round(c["library_setup_s"],2))
PY
