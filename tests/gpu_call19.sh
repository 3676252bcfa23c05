#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout -s KILL 120 python tests/gpu_diag.py gemm gemmperf > gpurun_out/diag_gemm4.log 2>&1; echo "gemm exit $?"; grep -E "gemmperf|torch|qkv-split|rc=|nan=[1-9]" gpurun_out/diag_gemm4.log | tail -14
timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --micro-batch 32 > gpurun_out/bench_mb32.log 2> gpurun_out/bench_mb32.err; echo "bench exit $?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_mb32.log",):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), d["clocks"], {k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
PY
