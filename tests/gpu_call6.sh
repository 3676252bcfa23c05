#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout 150 python tests/gpu_diag.py gemm gemmperf > gpurun_out/diag_gemm3.log 2>&1; echo "gemm exit $?"; grep -E "gemmperf|qkv-split|epi=2|epi=3" gpurun_out/diag_gemm3.log | tail -14
for t in attn; do
  ARIES_ATTN_POLY=0 timeout 200 python tests/prof_target.py $t > gpurun_out/plain_$t.log 2>&1 && \
  ARIES_ATTN_POLY=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_fwd" -s 2 -c 1 -o gpurun_out/prof3_$t -f python tests/prof_target.py $t > gpurun_out/ncu3_$t.log 2>&1
  echo "ncu $t exit $?"
done
