#!/bin/bash
# First GPU call of round 2 (about 6 minutes of box time): re-establish the baseline on the pool's B200 before any
# kernel work -- full GPU suite, smoke, default bench (N=1), the kernels next to the path, and an ncu launch list +
# one full capture of each kernel that round 2 is going to touch (CMC expand, the strip momentum kernel).
#   gpurun --timeout 900 -- 'bash scripts/gpu_round2_first.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_gpu.log
tail -3 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; cut -c1-300 gpurun_out/r2_bench_n1.json
timeout 300 python scripts/bench_cmc.py 128 > gpurun_out/r2_bench_cmc_128.json 2> gpurun_out/r2_bench_cmc_128.err; cat gpurun_out/r2_bench_cmc_128.json
timeout 300 python scripts/bench_configs.py > gpurun_out/r2_bench_configs.json 2> gpurun_out/r2_bench_configs.err; tail -c 400 gpurun_out/r2_bench_configs.json
# launch list of the CMC script (expand + transpose kernels) and a full capture of the expand kernel
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct \
  --clock-control none -k regex:cmc_ -c 12 --csv --log-file gpurun_out/r2_launches_cmc.csv python scripts/bench_cmc.py 96 > gpurun_out/r2_ncu_cmc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cmc_expand -s 2 -c 1 -o gpurun_out/r2_prof_cmc_expand -f \
  python scripts/bench_cmc.py 96 > gpurun_out/r2_ncu_cmc_full.log 2>&1
# the momentum kernel at 96^3: the capture every strip experiment of round 2 is compared with
timeout 600 ncu --set full --clock-control none --import-source on -k regex:staged_momentum -s 2 -c 1 -o gpurun_out/r2_prof_strip_mom -f \
  python bench.py --cells 96 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_mom.log 2>&1
ls -la gpurun_out | grep r2_ | head -20
