#!/bin/bash
# Round 2, GPU call 25 (2 GPUs): the default bench line at N = 2 (with the e2e leg and the parity check), as the driver's scaling run launches it.
mkdir -p gpurun_out
SECONDS=0
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29525 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_25_bench_n2.json 2> gpurun_out/r2_25_bench_n2.err; echo "bench exit $? wall ${SECONDS}s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_25_bench_n2.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["scaling"], "parity", d["multi_gpu_parity_max_rel_err"], "e2e", d["e2e"], "launches", d["gpu_launches"], "clocks", d["clocks"])
PY
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29526 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_25_ref_n2.json 2> gpurun_out/r2_25_ref_n2.err; echo "ref exit $? wall ${SECONDS}s"; cut -c1-200 gpurun_out/r2_25_ref_n2.json
