"""Same-box A/B timing of the encoder attention kernel alone (64 windows x 20 heads x 1500 tokens, the benchmarked launch):
    python tests/attn_ab.py            # from the tree whose library is to be timed (e.g. a worktree of the previous commit)
CUDA events, 3 warm-ups, 10 timed launches, best and mean."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from whisper_aries_b200 import _lib

B, T, H = 64, 1500, 20
d = 64 * H
dev = torch.device("cuda:0")
ctx = _lib.Context.get(0)
g = torch.Generator(device="cpu").manual_seed(1)
qk = (torch.randn(B * T, 2 * d, generator=g) * 1.0).to(dev).bfloat16().contiguous()
t_pad = (T + 7) // 8 * 8
vt = torch.randn(B, H, 64, t_pad, generator=g).to(dev).bfloat16().contiguous()
out = torch.empty(B * T, d, device=dev, dtype=torch.bfloat16)


def run():
    rc = ctx.lib.aries_test_attention(ctx.handle, qk.data_ptr(), vt.data_ptr(), B, T, H, t_pad, out.data_ptr(), None)
    assert rc == 0


for _ in range(3):
    run()
torch.cuda.synchronize()
times = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run()
    b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
print(f"{os.path.basename(ROOT)}: attention B={B}: best {min(times):.3f} ms, mean {sum(times) / len(times):.3f} ms, "
      f"checksum {out.float().abs().mean().item():.6f}")
