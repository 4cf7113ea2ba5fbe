#!/bin/bash
# Round 2, GPU call 38 (1 GPU): the default bench line with the per-kernel events inside the timed region.
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r2_38_bench_n1.json 2> gpurun_out/r2_38_bench_n1.err; echo "bench exit $? wall ${SECONDS}s"; grep "timed alone" gpurun_out/r2_38_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_38_bench_n1.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "sep", d["separate_kernels_ms_rank0"], "alone", d["kernels_timed_alone_ms_rank0"], "fused", d["fused_kernel_ms_rank0"], "frac", d["roofline"]["frac"], d["roofline"]["tracer"]["frac"], d["roofline"]["combined_frac"], "mom", d["roofline"]["kernel_ms"], "tracer", d["roofline"]["tracer"]["kernel_ms"], "e2e", d["e2e"]["value"], "setup", d["setup_s"], "traffic", d["roofline"]["traffic"])
PY
