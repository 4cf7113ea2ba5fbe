#!/bin/bash
# Round 2, GPU call 71 (2 GPUs, final library): the driver's own N = 2 command (strong scaling of the one 100 M-tet mesh, 1x1x2 node blocks,
# e2e leg and oracle parity check included).
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_71_bench_n2.json 2> gpurun_out/r2_71_bench_n2.err; echo "bench exit $? wall ${SECONDS}s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_71_bench_n2.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["scaling"], "parity", d["multi_gpu_parity_max_rel_err"], "halo", d["halo_update_ms_rank0"], "setup", d["setup_s"], "e2e", d["e2e"])
for r in d["per_rank"]: print(r["rank"], round(r["step_ms"],4), round(r["momentum_ms"],4), round(r["tracer_ms"],4), round(r["fused_ms"],4), round(r["halo_ms"],4), r["local_elements"])
PY
tail -3 gpurun_out/r2_71_bench_n2.err
