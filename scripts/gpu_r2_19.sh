#!/bin/bash
# Round 2, GPU call 19 (2 GPUs): multi-rank parity tests and the strong-scaling line at N = 2 with the keyed plans.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multirank.py -m gpu -q > gpurun_out/r2_19_pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_19_pytest_multi.log
tail -4 gpurun_out/r2_19_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-e2e > gpurun_out/r2_19_bench_n2.json 2> gpurun_out/r2_19_bench_n2.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_19_bench_n2.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["scaling"], "parity", d["multi_gpu_parity_max_rel_err"], "halo", d["halo_update_ms_rank0"])
for r in d["per_rank"]: print(r["rank"], round(r["step_ms"],4), round(r["momentum_ms"],4), round(r["tracer_ms"],4), round(r["halo_ms"],4))
PY
