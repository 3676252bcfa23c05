#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout -s KILL 300 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 8 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.log | cut -c1-200; tail -n 5 gpurun_out/bench.err
ARIES_ATTN_POLY=1 timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_poly1.log 2> gpurun_out/bench_poly1.err; echo "bench poly exit $?"
python - <<'PY'
import json
for f in ("gpurun_out/bench.log","gpurun_out/bench_poly1.log"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["ms_per_step"],2), d["clocks"], {k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
PY
