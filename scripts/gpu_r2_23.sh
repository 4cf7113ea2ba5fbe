#!/bin/bash
# Round 2, GPU call 23 (1 GPU): full suite with the free-surface stabilisation; smoke(); wall time of the default bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_23_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_23_pytest.log
tail -6 gpurun_out/r2_23_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2_23_smoke.log
SECONDS=0; timeout 1500 python bench.py > gpurun_out/r2_23_bench_n1.json 2> gpurun_out/r2_23_bench_n1.err; echo "bench exit $?"; echo "default bench wall ${SECONDS}s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_23_bench_n1.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "cpu", d["cpu_baseline"], "e2e", d["e2e"]["value"], "setup", d["setup_s"])
PY
SECONDS=0; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_23_bench_ref.json 2> gpurun_out/r2_23_bench_ref.err; cut -c1-400 gpurun_out/r2_23_bench_ref.json; echo "reference arm wall ${SECONDS}s"
