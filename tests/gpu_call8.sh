#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout 150 python tests/gpu_diag.py attn > gpurun_out/diag_attn.log 2>&1; echo "attn exit $?"; grep -E "attn" gpurun_out/diag_attn.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.log | cut -c1-300; tail -n 5 gpurun_out/bench.err
timeout 200 python tests/gpu_diag.py melperf > gpurun_out/diag_melperf.log 2>&1; cat gpurun_out/diag_melperf.log
