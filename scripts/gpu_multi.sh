#!/bin/bash
# usage: gpu_multi.sh N "cells list" [runtest]
N=${1:-2}
CELLS=${2:-"128 256"}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_multi.txt
if [ "${3:-1}" = "1" ]; then
timeout 600 python -m pytest tests/test_multirank.py -x -q -m gpu > gpurun_out/pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multi.log
tail -3 gpurun_out/pytest_multi.log
fi
for c in $CELLS; do
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --cells $c > gpurun_out/bench_multi${N}_$c.json 2> gpurun_out/bench_multi${N}_$c.err
tail -2 gpurun_out/bench_multi${N}_$c.err; cut -c1-200 gpurun_out/bench_multi${N}_$c.json
done
