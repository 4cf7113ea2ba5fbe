#!/bin/bash
# Round-1 evidence for profiles/: launch list of the bench command (all kernels), S3 launch list of the two
# STRIP kernels with DRAM bytes and pipe utilisation (source of roofline.traffic), --set full captures of
# both kernels at 96^3, smoke, reference arm.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_S3.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench_S3.log 2>&1
tail -1 gpurun_out/ncu_bench_S3.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct \
  --clock-control none -k regex:staged_ -s 2 -c 4 --csv --log-file gpurun_out/launches_S3_strip.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_S3_strip.log 2>&1
tail -1 gpurun_out/ncu_S3_strip.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:staged_momentum -s 2 -c 1 -o gpurun_out/prof_r1_strip_mom -f \
  python bench.py --cells 96 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_r1_mom.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:staged_advdiff -s 2 -c 1 -o gpurun_out/prof_r1_strip_adv -f \
  python bench.py --cells 96 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_r1_adv.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -2 gpurun_out/bench_reference.err
cut -c1-400 gpurun_out/bench_reference.json
