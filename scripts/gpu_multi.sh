#!/bin/bash
# usage: gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_multi.txt
timeout 600 python -m pytest tests/test_multirank.py -x -q -m gpu > gpurun_out/pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multi.log
tail -5 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --cells 128 --no-e2e > gpurun_out/bench_multi${N}_128.json 2> gpurun_out/bench_multi${N}_128.err
tail -3 gpurun_out/bench_multi${N}_128.err; cat gpurun_out/bench_multi${N}_128.json | cut -c1-400
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_multi${N}_256.json 2> gpurun_out/bench_multi${N}_256.err
tail -3 gpurun_out/bench_multi${N}_256.err; cat gpurun_out/bench_multi${N}_256.json | cut -c1-400
