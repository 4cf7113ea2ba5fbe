#!/bin/bash
# Round 2, GPU call 8 (1 GPU): parity + bench after the prologue/epilogue restructuring (batched loads, row table, circular plan queue, L2 prefetch of the successor block's metadata).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_8_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_8_pytest.log
tail -4 gpurun_out/r2_8_pytest.log
timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/r2_8_bench_n1.json 2> gpurun_out/r2_8_bench_n1.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_8_bench_n1.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "sep", d["separate_kernels_ms_rank0"], "fused", d["fused_kernel_ms_rank0"], "frac", d["roofline"]["frac"], d["roofline"]["tracer"]["frac"], "mom", d["roofline"]["kernel_ms"], "tracer", d["roofline"]["tracer"]["kernel_ms"])
for c in d["configs"]: print(c["config"][:44], round(c["momentum_ms"],3), round(c["tracer_ms"],3), round(c["gel_s"],2))
PY
