#!/bin/bash
# pytest -m gpu, smoke, bench, then ncu launch list + full captures of the three hot kernels.
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err
CMD="python bench.py --steps 1 --warmup 1 --windows 16 --no-cpu-baseline --no-e2e"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:logmel_tiles -c 1 -o gpurun_out/prof_mel -f $CMD > gpurun_out/ncu_mel.log 2>&1; echo "ncu mel exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 2 -c 1 -o gpurun_out/prof_attn -f $CMD > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 10 -c 4 -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"
ls -la gpurun_out
