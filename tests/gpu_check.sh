#!/bin/bash
# One standard GPU session (run under gpurun): per-kernel diagnostics, the GPU test-suite, the bench line.
# Every python step runs under `timeout -s KILL` so that a hung kernel cannot eat the box's time limit.
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
for t in mel gemm ln attn; do
  timeout -s KILL 150 python tests/gpu_diag.py $t > gpurun_out/diag_$t.log 2>&1; echo "diag $t exit $?"
done
grep -E "max_err|rror" gpurun_out/diag_attn.log | tail -4
timeout -s KILL 150 python tests/gpu_diag.py gemmperf melperf > gpurun_out/diag_perf.log 2>&1; grep -E "gemmperf|melperf|torch" gpurun_out/diag_perf.log
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 6 gpurun_out/pytest_gpu.log
timeout -s KILL 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench.log
timeout -s KILL 400 python bench.py --mode decode --steps 3 --warmup 1 > gpurun_out/bench_decode.log 2> gpurun_out/bench_decode.err; echo "bench decode exit $?"; cut -c1-300 gpurun_out/bench_decode.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
