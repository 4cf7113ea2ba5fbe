#!/bin/bash
# Round 2, GPU call 59 (1 GPU): the full GPU suite in a loop (no -x) to catch the rare surface-test failure with its details.
mkdir -p gpurun_out
for i in $(seq 1 8); do
  timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_59_run.log 2>&1
  if grep -q " failed" gpurun_out/r2_59_run.log; then echo "run $i FAILED: $(tail -1 gpurun_out/r2_59_run.log)"; grep -E "^FAILED" gpurun_out/r2_59_run.log | cut -c1-200 | head -8; cp gpurun_out/r2_59_run.log gpurun_out/r2_59_fail_$i.log; else echo "run $i ok: $(tail -1 gpurun_out/r2_59_run.log)"; fi
done | tee gpurun_out/r2_59_repeat.txt
