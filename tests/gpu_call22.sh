#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout -s KILL 90 python tests/gpu_diag.py attn > gpurun_out/diag_attn.log 2>&1; rc=$?; echo "attn exit $rc"; grep -E "attn|rror" gpurun_out/diag_attn.log | head -30
