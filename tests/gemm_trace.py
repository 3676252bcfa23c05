"""Timeline of the tcgen05 GEMM (run under gpurun; test infrastructure): where CTA 0's MMA-issue warp and first
epilogue warp spend their cycles per output tile.

    python tests/gemm_trace.py gemm-fc1|gemm-qkv|gemm-o|gemm-fc2

Uses tests/prof_target.py's shapes; the library reads the device buffer address from ARIES_GEMM_TRACE."""
import os
import runpy
import sys

import numpy as np
import torch

what = sys.argv[1] if len(sys.argv) > 1 else "gemm-fc1"
buf = torch.zeros(48 * 8, dtype=torch.int64, device="cuda")
os.environ["ARIES_GEMM_TRACE"] = str(buf.data_ptr())
sys.argv = ["prof_target.py", what]
runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "prof_target.py"), run_name="__main__")
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(48, 8).astype(np.int64)
n = int((t[:, 3] > 0).sum())
print(f"{what}: {n} traced tiles of CTA 0")
sl = slice(3, n - 1)
acc_wait = (t[:, 1] - t[:, 0])[sl]
loop = (t[:, 3] - t[:, 1])[sl]
op_wait = t[:, 2][sl]
period = np.diff(t[:n, 3])[2:-1]
print(f"  MMA warp per tile: accumulator-free wait {acc_wait.mean():.0f}  main loop {loop.mean():.0f} (of which waiting for operands "
      f"{op_wait.mean():.0f})  tile period {period.mean():.0f} cycles")
e_wait = (t[:, 5] - t[:, 4])[sl]
e_work = (t[:, 6] - t[:, 5])[sl]
print(f"  epilogue warp 0 per tile: waiting for the accumulator {e_wait.mean():.0f}  draining {e_work.mean():.0f} cycles")
