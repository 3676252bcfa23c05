"""Small profiling targets for ncu (run under gpurun):
    python tests/prof_target.py mel|mel-fused|attn|ln|gemm-o|gemm-fc1|gemm-fc2|gemm-qkv|dec-xattn|dec-skinny|dec-logits"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from whisper_aries_b200 import FeatureExtractor, _lib, synthetic   # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1]
ctx = _lib.Context.get(0)
lib = ctx.lib


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


if what == "mel":
    B = 64
    fe = FeatureExtractor(feature_size=128)
    xs = torch.from_numpy(synthetic.batch_signals(6, 0)).to(dev).repeat(11, 1)[:B].contiguous()
    out = torch.empty((B, 128, 3000), device=dev)
    for _ in range(3):
        fe(xs, frames_out=3000)
    torch.cuda.synchronize()
elif what == "mel-fused":
    # the fused PCM -> conv1-operand launch of aries_encode_pcm (128 mel bins) in front of a toy encoder
    from whisper_aries_b200 import WhisperModel
    B = 64
    shape = synthetic.EncoderShape("mel128-toy", 128, 128, 2, 2, 512)
    model = WhisperModel(shape, synthetic.encoder_weights(shape, 1), device="cuda", device_index=0)
    xs = torch.from_numpy(synthetic.batch_signals(6, 0)).to(dev).repeat(11, 1)[:B].contiguous()
    for _ in range(3):
        model.encode_audio(xs)
    torch.cuda.synchronize()
elif what == "attn":
    B_, T_, H = 8, 1500, 20
    d, t_pad = 64 * H, 1504
    qk = (torch.randn(B_ * T_, 2 * d, device=dev) * 1.5).bfloat16()
    vt = torch.randn(B_, H, 64, t_pad, device=dev).bfloat16()
    out = torch.empty((B_ * T_, d), device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        lib.aries_test_attention(ctx.handle, ptr(qk), ptr(vt), B_, T_, H, t_pad, ptr(out), None)
    torch.cuda.synchronize()
elif what == "ln":
    rows, d = 96000, 1280
    x = torch.randn(rows, d, device=dev).half()
    gm, bt = torch.randn(d, device=dev), torch.randn(d, device=dev)
    y = torch.empty((rows, d), device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        lib.aries_test_layernorm(ctx.handle, ptr(x), ptr(gm), ptr(bt), ptr(y), rows, d, None)
    torch.cuda.synchronize()
elif what == "dec-xattn":
    # cross-attention of one decode step at the bench shape: 64 windows x 20 heads x 1500 keys (491.5 MB of keys / values)
    B_, H, n = 64, 20, 1500
    d = 64 * H
    q = torch.randn(B_, d, device=dev).bfloat16()
    kv = torch.randn(B_ * n, 2 * d, device=dev).bfloat16()
    out = torch.empty((B_, d), device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        lib.aries_test_decode_attention(ctx.handle, ptr(q), d, ptr(kv), ctypes.c_void_p(kv.data_ptr() + d * 2), n, 2 * d,
                                        None, None, 0, None, n, B_, H, ptr(out), d, 1, None)
    torch.cuda.synchronize()
elif what in ("dec-skinny", "dec-logits"):
    # the fc1 projection (13.1 MB of weights) / the logits projection (132.8 MB) of one decode step, 64 sequences
    B_, N, K, epi = (64, 5120, 1280, 1) if what == "dec-skinny" else (64, 51866, 1280, 3)
    x = torch.randn(B_, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.empty((B_, N), device=dev, dtype=torch.float32 if epi == 3 else torch.bfloat16)
    for _ in range(3):
        lib.aries_test_skinny_gemm(ctx.handle, epi, B_, N, K, ptr(x), ptr(w), ptr(bias), ptr(out), N, 0, None)
    torch.cuda.synchronize()
else:
    # the GEMMs of one encoder layer at the bench shape (M = 64 x 1500) with the epilogues the encoder really uses:
    #   gemm-fc1  LayerNorm-folded fc1 + GELU (epi 5)            gemm-qkv  LayerNorm-folded QKV, Q|K row-major, V transposed (epi 6)
    #   gemm-o / gemm-fc2  bias + f16 residual, LayerNorm partials written (epi 2 with stats_out)
    M, T_ = 96000, 1500
    N, K, epi = {"gemm-o": (1280, 1280, 2), "gemm-fc1": (5120, 1280, 5), "gemm-qkv": (3840, 1280, 6),
                 "gemm-fc2": (1280, 5120, 2)}[what]
    bias = torch.randn(N, device=dev)
    if epi == 2:
        a = torch.randn(M, K, device=dev).bfloat16()
        b = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        resid = torch.randn(M, N, device=dev).half()
        out = torch.empty((M, N), device=dev, dtype=torch.float16)
        parts = lib.aries_test_gemm_stats_parts(N)
        stats = torch.empty((M, parts, 2), device=dev)

        def run():
            lib.aries_test_gemm_ln(ctx.handle, 2, M, N, K, ptr(a), ptr(b), ptr(bias), None, None, 0, 0, ptr(resid), ptr(out),
                                   None, 0, 0, 0, ptr(stats), None)
    else:
        a = (torch.randn(M, K, device=dev) * 2 + 0.3).half()                    # the f16 residual stream
        b = (torch.randn(N, K, device=dev) * 0.05).half()                       # gamma-scaled f16 weights
        c1 = b.float().sum(1).contiguous()
        parts = lib.aries_test_gemm_stats_parts(K)
        xs = a.float().view(M, parts, K // parts)
        stats = torch.stack([xs.sum(2), (xs * xs).sum(2)], dim=2).contiguous()
        del xs
        if epi == 5:
            out = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
            out2, n_split, t_pad = None, 0, 0
        else:
            out = torch.empty((M, 2 * K), device=dev, dtype=torch.bfloat16)
            out2 = torch.empty((M // T_, K // 64, 64, 1504), device=dev, dtype=torch.bfloat16)
            n_split, t_pad = 2 * K, 1504

        def run():
            lib.aries_test_gemm_ln(ctx.handle, epi, M, N, K, ptr(a), ptr(b), ptr(bias), ptr(c1), ptr(stats), parts, K, None,
                                   ptr(out), ptr(out2), n_split, T_, t_pad, None, None)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
print("ok", what)
