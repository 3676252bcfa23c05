#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout -s KILL 90 python tests/gpu_diag.py attn > gpurun_out/diag_attn.log 2>&1; rc=$?; echo "attn exit $rc"; grep -E "attn|rror" gpurun_out/diag_attn.log | head -30
if [ $rc -ne 0 ]; then exit 1; fi
ARIES_ATTN_TOKEN=1 timeout -s KILL 90 python tests/gpu_diag.py attn > gpurun_out/diag_attn_token.log 2>&1; echo "attn token exit $?"; grep -E "attn time" gpurun_out/diag_attn_token.log | head -30
ARIES_ATTN_POLY=1 timeout -s KILL 90 python tests/gpu_diag.py attn > gpurun_out/diag_attn_poly1.log 2>&1; echo "attn poly exit $?"; grep -E "attn time" gpurun_out/diag_attn_poly1.log | head -30
timeout -s KILL 100 python tests/prof_target.py attn > gpurun_out/plain_attn.log 2>&1 && \
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"attention_fwd" -s 2 -c 1 -o gpurun_out/prof8_attn -f python tests/prof_target.py attn > gpurun_out/ncu8_attn.log 2>&1
echo "ncu attn exit $?"
