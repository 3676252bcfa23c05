#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
for w in 8 16 32; do
timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --windows $w --no-cpu-baseline --no-e2e > gpurun_out/bench_w$w.log 2> gpurun_out/bench_w$w.err; echo "bench w$w exit $?"
done
python - <<'PY'
import json
for w in (8,16,32):
    d=json.loads(open(f"gpurun_out/bench_w{w}.log").read().strip().splitlines()[-1])
    print(w, round(d["value"]), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms_per_step"]*64/w,2) for k,v in d["kernels"].items()})
PY
