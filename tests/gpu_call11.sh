#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout -s KILL 90 python tests/gpu_diag.py attn > gpurun_out/diag_attn.log 2>&1; rc=$?; echo "attn exit $rc"; grep -E "attn|rror" gpurun_out/diag_attn.log | head -30
if [ $rc -ne 0 ]; then exit 1; fi
ARIES_ATTN_NOTOKEN=1 timeout -s KILL 90 python tests/gpu_diag.py attn > gpurun_out/diag_attn_notoken.log 2>&1; echo "attn notoken exit $?"; grep -E "attn time" gpurun_out/diag_attn_notoken.log | head -30
ARIES_ATTN_POLY=1 timeout -s KILL 90 python tests/gpu_diag.py attn > gpurun_out/diag_attn_poly1.log 2>&1; echo "attn poly exit $?"; grep -E "attn time" gpurun_out/diag_attn_poly1.log | head -30
timeout -s KILL 300 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 8 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.log | cut -c1-300; tail -n 5 gpurun_out/bench.err
timeout -s KILL 100 python tests/prof_target.py attn > gpurun_out/plain_attn.log 2>&1 && \
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"attention_fwd" -s 2 -c 1 -o gpurun_out/prof6_attn -f python tests/prof_target.py attn > gpurun_out/ncu6_attn.log 2>&1
echo "ncu attn exit $?"
