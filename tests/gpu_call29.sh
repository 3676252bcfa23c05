#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout -s KILL 120 python tests/gpu_diag.py ln > gpurun_out/diag_ln.log 2>&1; grep -E "ln d=" gpurun_out/diag_ln.log
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
for f in ("gpurun_out/bench.log",):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], {k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
PY
