#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_momentum_walk -s 2 -c 1 -o gpurun_out/prof_walk_mom -f \
  python bench.py --cells 96 --scatter gather --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_direct.log 2>&1
tail -2 gpurun_out/ncu_direct.log
