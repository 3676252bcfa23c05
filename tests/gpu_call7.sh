#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
ARIES_ATTN_POLY=0 timeout 150 python tests/gpu_diag.py attn > gpurun_out/diag_attn_poly0.log 2>&1; echo "attn poly0 exit $?"; grep -E "attn" gpurun_out/diag_attn_poly0.log
ARIES_ATTN_POLY=1 timeout 150 python tests/gpu_diag.py attn > gpurun_out/diag_attn_poly1.log 2>&1; echo "attn poly1 exit $?"; grep -E "attn time" gpurun_out/diag_attn_poly1.log
timeout 200 python tests/gpu_diag.py enc-tiny > gpurun_out/diag_enc-tiny.log 2>&1; grep "enc tiny" gpurun_out/diag_enc-tiny.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.log | cut -c1-300; tail -n 5 gpurun_out/bench.err
ARIES_ATTN_POLY=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_poly1.log 2> gpurun_out/bench_poly1.err; echo "bench poly1 exit $?"
ARIES_ATTN_POLY=0 timeout 200 python tests/prof_target.py attn > gpurun_out/plain_attn.log 2>&1 && \
ARIES_ATTN_POLY=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_fwd" -s 2 -c 1 -o gpurun_out/prof4_attn -f python tests/prof_target.py attn > gpurun_out/ncu4_attn.log 2>&1
echo "ncu attn exit $?"
