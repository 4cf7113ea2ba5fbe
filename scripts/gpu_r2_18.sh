#!/bin/bash
# Round 2, GPU call 18 (1 GPU): geometric keys in the plan builder (renumbered meshes), full suite, bench with configs.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_18_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_18_pytest.log
tail -6 gpurun_out/r2_18_pytest.log
timeout 1500 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2_18_bench_n1.json 2> gpurun_out/r2_18_bench_n1.err; echo "bench exit $?"; tail -3 gpurun_out/r2_18_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_18_bench_n1.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "sep", d["separate_kernels_ms_rank0"], "fused", d["fused_kernel_ms_rank0"], "setup", d["setup_s"])
for c in d["configs"]:
    if "error" in c: print(c); continue
    print(c["config"][:60], round(c["momentum_ms"],3), round(c["tracer_ms"],3), round(c["gel_s"],2), c["momentum_path"], c["tracer_path"], round(c["library_setup_s"],1))
PY
