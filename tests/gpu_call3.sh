#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
run() { timeout "$2" python tests/gpu_diag.py $1 > "gpurun_out/diag_$1.log" 2>&1; rc=$?; echo "$1 exit $rc"; tail -n 45 "gpurun_out/diag_$1.log"; return $rc; }
run mel 200
run melperf 150
run gemm 150 && run gemmperf 150
run attn 150
run enc-tiny 200
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err
