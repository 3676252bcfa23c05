#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
ARIES_GEMM_PAIR=0 timeout -s KILL 120 python tests/gpu_diag.py gemm > gpurun_out/diag_gemm_p0.log 2>&1; echo "gemm pair0 exit $?"; grep -E "rror|rc=[^0]|nan=[1-9]" gpurun_out/diag_gemm_p0.log | head; grep -c "max_err" gpurun_out/diag_gemm_p0.log
ARIES_GEMM_PAIR=1 timeout -s KILL 120 python tests/gpu_diag.py gemm gemmperf > gpurun_out/diag_gemm_p1.log 2>&1; echo "gemm pair1 exit $?"; grep -E "gemm " gpurun_out/diag_gemm_p1.log | awk '{print $2,$3,$4,$5,$6,$7,$8}' | head -40; grep -E "gemmperf|torch|rror" gpurun_out/diag_gemm_p1.log | tail -12
