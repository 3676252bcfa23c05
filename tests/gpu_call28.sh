#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout -s KILL 120 python tests/gpu_diag.py gemm gemmperf ln > gpurun_out/diag_gemm6.log 2>&1; echo "gemm exit $?"; grep -E "epi=2|epi=3|ln d=" gpurun_out/diag_gemm6.log | awk '{print $2,$3,$4,$5,$6,$7,$8}' | head -20; grep -E "gemmperf|rror" gpurun_out/diag_gemm6.log | tail -6
timeout -s KILL 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 6 gpurun_out/pytest_gpu.log
timeout -s KILL 200 python tests/gpu_diag.py enc-tiny > gpurun_out/diag_enc-tiny.log 2>&1; grep "enc tiny: vs" gpurun_out/diag_enc-tiny.log
timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
for f in ("gpurun_out/bench.log",):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["ms_per_step"],2), "e2e", d["e2e"] and round(d["e2e"]["value"]), d["clocks"], "roof", round(d["roofline"]["frac"],3), {k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
PY
