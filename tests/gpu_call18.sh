#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
for pz in 0 8 4; do
ARIES_ATTN_POLY=$pz timeout -s KILL 90 python tests/gpu_diag.py attn > gpurun_out/diag_attn_poly$pz.log 2>&1; echo "attn poly $pz exit $?"; grep -E "attn time|max_err" gpurun_out/diag_attn_poly$pz.log | tail -3
done
for pz in 0 4; do
ARIES_ATTN_POLY=$pz timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_poly$pz.log 2> gpurun_out/bench_poly$pz.err; echo "bench poly $pz exit $?"
done
python - <<'PY'
import json
for f in ("gpurun_out/bench_poly0.log","gpurun_out/bench_poly4.log"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["ms_per_step"],2), d["clocks"], {k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
PY
