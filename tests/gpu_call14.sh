#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
ARIES_ATTN_TRACE=1 timeout -s KILL 90 python tests/attn_trace.py 8 > gpurun_out/attn_trace.log 2>&1; echo "trace exit $?"; cat gpurun_out/attn_trace.log | head -60
