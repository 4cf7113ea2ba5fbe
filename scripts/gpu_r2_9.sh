#!/bin/bash
# Round 2, GPU call 9 (1 GPU): parity of the new entry points (mass matrix, continuity by parts, vector Dirichlet, velocity correction), fused write-out fix.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_9_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_9_pytest.log
tail -4 gpurun_out/r2_9_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-configs --no-e2e > gpurun_out/r2_9_bench_n1.json 2> gpurun_out/r2_9_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_9_bench_n1.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "sep", d["separate_kernels_ms_rank0"], "fused", d["fused_kernel_ms_rank0"])
PY
