#!/bin/bash
# First GPU contact: per-kernel diagnostics, each isolated under its own timeout.
mkdir -p gpurun_out
python -c "import torch; print(torch.__version__, torch.cuda.is_available())" > gpurun_out/warm.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
run() { # name timeout
  timeout "$2" python tests/gpu_diag.py $1 > "gpurun_out/diag_$1.log" 2>&1
  rc=$?
  echo "$1 exit $rc"
  tail -n 40 "gpurun_out/diag_$1.log"
  return $rc
}
run ln 150
run mel 240
run melperf 150
if run gemm 150; then
  run gemmperf 150
  run attn 150
  run enc-micro 200 && run enc-tiny 200
fi
