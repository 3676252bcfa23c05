#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout 150 python tests/gpu_diag.py gemm gemmperf > gpurun_out/diag_gemm2.log 2>&1; echo "gemm exit $?"; grep -E "gemmperf|qkv-split" gpurun_out/diag_gemm2.log
for t in mel attn gemm-o gemm-qkv; do
  timeout 200 python tests/prof_target.py $t > gpurun_out/plain_$t.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"logmel_tiles|attention_fwd|gemm_bf16" -s 2 -c 1 -o gpurun_out/prof2_$t -f python tests/prof_target.py $t > gpurun_out/ncu2_$t.log 2>&1
  echo "ncu $t exit $?"
done
