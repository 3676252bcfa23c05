#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout 200 python tests/prof_target.py attn > gpurun_out/plain_attn.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_fwd" -s 2 -c 1 -o gpurun_out/prof5_attn -f python tests/prof_target.py attn > gpurun_out/ncu5_attn.log 2>&1
echo "ncu attn exit $?"
