#!/bin/bash
# Round 2, GPU call 12 (1 GPU): full absorption matrix inside the common STRIP kernel (strip_absorb.cu), Delaunay
# unstructured meshes (parity + configs leg), SU configs leg, corrected shared-memory microbenchmark, pair-per-record staging.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "additive or delaunay or fused or abi or shuffled or large or strip" > gpurun_out/r2_12_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_12_pytest.log
tail -15 gpurun_out/r2_12_pytest.log
timeout 120 scripts/bin/microbench_fp64_lds > gpurun_out/r2_12_microbench.txt 2>&1; grep -A1 "LDS" gpurun_out/r2_12_microbench.txt | head -4
timeout 1500 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2_12_bench_n1.json 2> gpurun_out/r2_12_bench_n1.err; echo "bench exit $?"; tail -3 gpurun_out/r2_12_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_12_bench_n1.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "sep", d["separate_kernels_ms_rank0"], "fused", d["fused_kernel_ms_rank0"])
for c in d["configs"]:
    if "error" in c: print(c); continue
    print(c["config"][:60], round(c["momentum_ms"],3), round(c["tracer_ms"],3), round(c["gel_s"],2), c["momentum_path"], c["tracer_path"], round(c["library_setup_s"],1), c.get("plan"))
PY
