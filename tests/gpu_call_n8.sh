#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
nvidia-smi -L | wc -l
N=$(nvidia-smi -L | wc -l)
timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; echo "bench n$N exit $?"; tail -n 3 gpurun_out/bench_n8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n8.log").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"]), round(d["ms_per_step"],2), "e2e", d["e2e"] and round(d["e2e"]["value"]), d["clocks"])
PY
