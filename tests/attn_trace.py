"""Timeline of the attention kernel (run under gpurun with ARIES_ATTN_TRACE=1; test infrastructure).

    ARIES_ATTN_TRACE=1 python tests/attn_trace.py [batch]

Prints, for a few traced CTAs, the average clock64 cycles the softmax warps spend per key tile in each phase
(wait S | load+max | exponentials+pack | store+publish) and what the MMA-issue warps spend (wait P | PV issue | QK issue)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from whisper_aries_b200 import _lib   # noqa: E402

B_ = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T_, H = 1500, 20
dev = torch.device("cuda:0")
ctx = _lib.Context.get(0)
lib = ctx.lib
d, t_pad = 64 * H, 1504
qk = (torch.randn(B_ * T_, 2 * d, device=dev) * 1.5).bfloat16()
vt = torch.randn(B_, H, 64, t_pad, device=dev).bfloat16()
out = torch.empty((B_ * T_, d), device=dev, dtype=torch.bfloat16)
p = lambda t: ctypes.c_void_p(t.data_ptr())
for _ in range(3):
    lib.aries_test_attention(ctx.handle, p(qk), p(vt), B_, T_, H, t_pad, p(out), None)
torch.cuda.synchronize()
CT, W, E = 16, 11, 160
QTILES = int(os.environ.get('ARIES_ATTN_QTILES', '1'))          # must match the library build
SOFTMAX_WARPS = (0, 3) if QTILES == 1 else (0, 3, 4, 7)
MMA_WARPS = (5,) if QTILES == 1 else (9, 10)
buf = np.zeros(CT * W * E, dtype=np.uint64)
rc = lib.aries_test_attention_trace(ctx.handle, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
assert rc == 0
tr = buf.reshape(CT, W, E).astype(np.int64)
np.save("gpurun_out/attn_trace.npy", tr)
n_kv = 24
for c in range(CT):
    if tr[c, 0, 0] == 0:
        continue
    t0 = tr[c, :, 0].min()
    print(f"--- traced CTA slot {c}: lifetime {tr[c].max() - t0} clk")
    for w in SOFTMAX_WARPS:                       # softmax stamps: entry, 4 per tile, 2 (o_full), 2 (exit)
        s = tr[c, w]
        ev = s[1:1 + 4 * n_kv].reshape(n_kv, 4)
        nxt = np.concatenate([ev[1:, 0], s[1 + 4 * n_kv:2 + 4 * n_kv]])
        wait, ldmax, exp, stpub = ev[:, 1] - ev[:, 0], ev[:, 2] - ev[:, 1], ev[:, 3] - ev[:, 2], nxt - ev[:, 3]
        sl = slice(3, n_kv - 2)
        k = 1 + 4 * n_kv
        print(f"  softmax warp {w}: first S ready at +{ev[0, 1] - t0}; per tile (tiles 3..{n_kv - 3}) wait_S {wait[sl].mean():.0f} "
              f"ld+max {ldmax[sl].mean():.0f} exp+pack {exp[sl].mean():.0f} st+publish {stpub[sl].mean():.0f} "
              f"total {(ev[n_kv - 2, 0] - ev[3, 0]) / (n_kv - 5):.0f}; o_full wait {s[k + 1] - s[k]}; "
              f"epilogue {s[k + 2] - s[k + 1]}; exit at +{s[k + 3] - t0}")
    for w in MMA_WARPS:                            # MMA issue: 1 + 4 per tile
        s = tr[c, w]
        ev = s[1:1 + 4 * n_kv].reshape(n_kv, 4)
        sl = slice(3, n_kv - 3)
        print(f"  mma warp {w}: per tile wait_P {(ev[:, 1] - ev[:, 0])[sl].mean():.0f} v_full+PV issue "
              f"{(ev[:, 2] - ev[:, 1])[sl].mean():.0f} k_full+QK issue+commits {(ev[:, 3] - ev[:, 2])[sl].mean():.0f} "
              f"loop overhead {(ev[1:, 0] - ev[:-1, 3])[sl].mean():.0f}")
