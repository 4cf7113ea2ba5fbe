#!/bin/bash
# Round 2, GPU call 10 (1 GPU): state check after the permuted node-record mirrors (parity + bench with the configs block),
# FP64 / shared-memory microbenchmark (the per-SM ceilings of the STRIP kernels), launch list at 128^3.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_10_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_10_pytest.log
tail -3 gpurun_out/r2_10_pytest.log
timeout 120 scripts/bin/microbench_fp64_lds > gpurun_out/r2_10_microbench.txt 2>&1; cat gpurun_out/r2_10_microbench.txt
timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/r2_10_bench_n1.json 2> gpurun_out/r2_10_bench_n1.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_10_bench_n1.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "sep", d["separate_kernels_ms_rank0"], "fused", d["fused_kernel_ms_rank0"], "frac", d["roofline"]["frac"], d["roofline"]["tracer"]["frac"], "mom", d["roofline"]["kernel_ms"], "tracer", d["roofline"]["tracer"]["kernel_ms"], "e2e", d["e2e"])
for c in d["configs"]: print(c["config"][:60], round(c["momentum_ms"],3), round(c["tracer_ms"],3), round(c["gel_s"],2))
PY
