#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout 150 python tests/gpu_diag.py attn > gpurun_out/diag_attn.log 2>&1; echo "attn exit $?"; grep -E "attn|rror" gpurun_out/diag_attn.log | head -30
ARIES_ATTN_POLY=1 timeout 150 python tests/gpu_diag.py attn > gpurun_out/diag_attn_poly1.log 2>&1; echo "attn poly exit $?"; grep -E "attn time" gpurun_out/diag_attn_poly1.log | head -30
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.log | cut -c1-300; tail -n 5 gpurun_out/bench.err
