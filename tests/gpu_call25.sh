#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout -s KILL 90 python tests/gpu_diag.py attn > gpurun_out/diag_attn.log 2>&1; rc=$?; echo "attn exit $rc"; grep -E "attn|rror" gpurun_out/diag_attn.log | head -30
ARIES_ATTN_TRACE=1 timeout -s KILL 90 python tests/attn_trace.py 8 > gpurun_out/attn_trace.log 2>&1; echo "trace exit $?"; cat gpurun_out/attn_trace.log | head -8
