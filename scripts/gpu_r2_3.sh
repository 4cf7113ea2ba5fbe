#!/bin/bash
# Round 2, GPU call 3 (2 GPUs): additive-pass parity, the 2-GPU NCCL test (plain + overlapped exchange), strong-scaling bench at N=2.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_3_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_3_pytest.log
tail -5 gpurun_out/r2_3_pytest.log
CGASM_DEBUG=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_3_bench_n2.json 2> gpurun_out/r2_3_bench_n2.err; echo "bench n2 exit $?"; cut -c1-400 gpurun_out/r2_3_bench_n2.json; tail -5 gpurun_out/r2_3_bench_n2.err
CGASM_DEBUG=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-overlap --no-e2e --no-parity > gpurun_out/r2_3_bench_n2_noov.json 2> gpurun_out/r2_3_bench_n2_noov.err; echo "bench n2 noov exit $?"; cut -c1-200 gpurun_out/r2_3_bench_n2_noov.json
