#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
nvidia-smi -L
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"; tail -c 1500 gpurun_out/bench_n2.log | cut -c1-400; tail -n 5 gpurun_out/bench_n2.err
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2.log 2> gpurun_out/bench_ref_n2.err; echo "ref n2 exit $?"; cat gpurun_out/bench_ref_n2.log | cut -c1-600
