"""GPU: each encoder kernel on its own (test hooks of the C ABI) against a plain torch fp32 reference of the same op.
Tolerances: bf16 outputs within 2^-8 relative of the row scale; f32 outputs within bf16-operand rounding."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


@pytest.fixture(scope="module")
def ctx():
    from whisper_aries_b200 import _lib
    return _lib.Context.get(0)


def gemm_ref(a, b, bias, epi, resid, pos, pos_rows):
    acc = a.float() @ b.float().t() + bias
    if epi in (1, 3):
        acc = torch.nn.functional.gelu(acc)
    if epi == 2:
        acc = acc + resid
    if epi == 3:
        acc = acc + pos[torch.arange(a.shape[0], device=a.device) % pos_rows]
    return acc


@pytest.mark.parametrize("shape", [(128, 128, 64), (300, 384, 128), (1500, 1280, 1280), (2900, 3840, 1280),
                                   (700, 1280, 5120), (1000, 512, 128)])
@pytest.mark.parametrize("epi", [0, 1, 2, 3])
def test_gemm_epilogues(ctx, shape, epi):
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N + K + epi)
    a = (torch.randn(M, K, generator=g) * 0.5).cuda().bfloat16()
    b = (torch.randn(N, K, generator=g) * 0.05).cuda().bfloat16()
    bias = torch.randn(N, generator=g).cuda()
    resid = torch.randn(M, N, generator=g).cuda().half()          # the residual stream is f16
    pos_rows = 100
    pos = torch.randn(pos_rows, N, generator=g).cuda()
    Mx = (M // pos_rows) * pos_rows if epi == 3 else M
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16 if epi < 2 else torch.float16)
    from whisper_aries_b200 import _lib
    _lib.check(ctx.lib.aries_test_gemm(ctx.handle, epi, Mx, N, K, ptr(a), ptr(b), ptr(bias), ptr(resid), ptr(pos),
                                       pos_rows, ptr(out), None, 0, 0, 0, None))
    torch.cuda.synchronize()
    ref = gemm_ref(a[:Mx], b, bias, epi, resid[:Mx].float(), pos, pos_rows)
    tol = 2e-3 + ref.abs().max().item() * (2 ** -8 if epi < 2 else 2 ** -11)      # bf16 / f16 output rounding
    assert (out[:Mx].float() - ref).abs().max().item() <= tol
    if Mx < M:
        assert torch.isnan(out[Mx:].float()).all()                 # rows past M are never written


def test_gemm_qkv_split(ctx):
    from whisper_aries_b200 import _lib
    B, T, d, K, t_pad = 2, 300, 256, 128, 304
    g = torch.Generator().manual_seed(5)
    a = (torch.randn(B * T, K, generator=g) * 0.5).cuda().bfloat16()
    b = (torch.randn(3 * d, K, generator=g) * 0.05).cuda().bfloat16()
    bias = torch.randn(3 * d, generator=g).cuda()
    out = torch.zeros((B * T, 2 * d), device="cuda", dtype=torch.bfloat16)
    out2 = torch.zeros((B, d // 64, 64, t_pad), device="cuda", dtype=torch.bfloat16)
    _lib.check(ctx.lib.aries_test_gemm(ctx.handle, 4, B * T, 3 * d, K, ptr(a), ptr(b), ptr(bias), None, None, 0,
                                       ptr(out), ptr(out2), 2 * d, T, t_pad, None))
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t() + bias
    assert (out.float() - ref[:, : 2 * d]).abs().max().item() <= 0.05
    v = ref[:, 2 * d:].reshape(B, T, d // 64, 64).permute(0, 2, 3, 1)
    assert (out2[..., :T].float() - v).abs().max().item() <= 0.05
    assert out2[..., T:].abs().max().item() == 0


def test_gemm_rejects_bad_shapes(ctx):
    from whisper_aries_b200 import _lib
    x = torch.zeros(128 * 128, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        _lib.check(ctx.lib.aries_test_gemm(ctx.handle, 0, 128, 100, 64, ptr(x), ptr(x), ptr(x), None, None, 0, ptr(x),
                                           None, 0, 0, 0, None))


@pytest.mark.parametrize("d", [128, 384, 1024, 1280])
def test_layernorm(ctx, d):
    from whisper_aries_b200 import _lib
    x = (torch.randn(999, d, device="cuda") * 3 + 0.5).half()           # the residual stream is f16
    gm, bt = torch.randn(d, device="cuda"), torch.randn(d, device="cuda")
    y = torch.empty((999, d), device="cuda", dtype=torch.bfloat16)
    _lib.check(ctx.lib.aries_test_layernorm(ctx.handle, ptr(x), ptr(gm), ptr(bt), ptr(y), 999, d, None))
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.float(), (d,), gm, bt, 1e-5)
    assert (y.float() - ref).abs().max().item() <= ref.abs().max().item() * 2 ** -8 + 1e-3


@pytest.mark.parametrize("cfg", [(1, 128, 1), (1, 200, 2), (2, 1500, 2), (1, 1500, 20)])
def test_attention(ctx, cfg):
    from whisper_aries_b200 import _lib
    B, T, H = cfg
    d, t_pad = 64 * H, (T + 7) // 8 * 8
    g = torch.Generator().manual_seed(T + H)
    q, k, v = (torch.randn(B, T, H, 64, generator=g).cuda() * s for s in (1.5, 1.5, 1.0))
    qk = torch.cat([q.reshape(B * T, d), k.reshape(B * T, d)], dim=1).bfloat16().contiguous()
    vt = torch.full((B, H, 64, t_pad), float("nan"), device="cuda", dtype=torch.bfloat16)   # pad must never be read
    vt[..., :T] = v.permute(0, 2, 3, 1).bfloat16()
    out = torch.full((B * T, d), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(ctx.lib.aries_test_attention(ctx.handle, ptr(qk), ptr(vt), B, T, H, t_pad, ptr(out), None))
    torch.cuda.synchronize()
    qf, kf, vf = (t.bfloat16().float().permute(0, 2, 1, 3) for t in (q, k, v))
    ref = (torch.softmax(qf @ kf.transpose(-1, -2) / 8.0, dim=-1) @ vf).permute(0, 2, 1, 3).reshape(B * T, d)
    assert (out.float() - ref).abs().max().item() <= 0.03          # P and the output are rounded to bf16


@pytest.mark.parametrize("n", [1, 7, 8, 4096 + 3, 480000, 64 * 480000])
def test_pcm_s16_to_f32_is_exact(n):
    """Row f3: the device-side int16 -> float32 / 32768 conversion equals numpy's, bit for bit (also ragged lengths and an
    unaligned view), at the bench size (64 windows)."""
    import numpy as np
    from whisper_aries_b200 import pcm_s16_to_f32
    rng = np.random.default_rng(n)
    x = rng.integers(-32768, 32768, size=n + 1, dtype=np.int16)
    x[:2] = [-32768, 32767]
    xd = torch.from_numpy(x).cuda()
    got = pcm_s16_to_f32(xd[:n])
    assert np.array_equal(got.cpu().numpy(), x[:n].astype(np.float32) / 32768.0)
    got = pcm_s16_to_f32(xd[1:])                 # 2-byte aligned only: scalar kernel
    assert np.array_equal(got.cpu().numpy(), x[1:].astype(np.float32) / 32768.0)


def test_scheduler_accepts_int16_windows():
    """gpu_worker with int16 PCM gives the same encoder states as with the float32 conversion done on the host."""
    import numpy as np
    from oracle import synth as osynth
    from whisper_aries_b200 import ChunkScheduler, WhisperModel, gpu_worker, synthetic
    shape = synthetic.SHAPES["micro"]
    model = WhisperModel(shape, synthetic.encoder_weights(shape, 1234), device="cuda", device_index=0)
    pcm16 = np.clip(np.round(osynth.batch_signals(5, 0) * 32768.0), -32768, 32767).astype(np.int16)
    out16 = torch.empty((5, shape.n_ctx, shape.d_model), dtype=torch.bfloat16).pin_memory()
    out32 = torch.empty_like(out16).pin_memory()
    sched = ChunkScheduler([gpu_worker(model, micro_batch=2)])
    assert all(r.success for r in sched.run(torch.from_numpy(pcm16).pin_memory(), out16))
    assert all(r.success for r in sched.run(torch.from_numpy(pcm16.astype(np.float32) / 32768.0).pin_memory(), out32))
    assert torch.equal(out16, out32)


@pytest.mark.parametrize("shape", [(300, 128, 128), (700, 384, 384), (1500, 1280, 1280), (2900, 1024, 4096)])
def test_residual_epilogue_leaves_layernorm_partials(ctx, shape):
    """Producing side of the LayerNorm fold: the residual epilogue (epi 2) also writes, per output row and per column
    slice, the sum and the sum of squares of what it stores; summed over the slices they must give the row's mean and
    variance of the f16 stream (f32 arithmetic: rel 1e-4 on the sums of squares)."""
    from whisper_aries_b200 import _lib
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N)
    a = (torch.randn(M, K, generator=g) * 0.5).cuda().bfloat16()
    b = (torch.randn(N, K, generator=g) * 0.05).cuda().bfloat16()
    bias = torch.randn(N, generator=g).cuda()
    resid = (torch.randn(M, N, generator=g) * 3 + 0.7).cuda().half()
    parts = ctx.lib.aries_test_gemm_stats_parts(N)
    assert parts == N // (128 if N % 256 == 0 else 64)
    out = torch.empty((M, N), device="cuda", dtype=torch.float16)
    stats = torch.full((M, parts, 2), float("nan"), device="cuda")
    _lib.check(ctx.lib.aries_test_gemm_ln(ctx.handle, 2, M, N, K, ptr(a), ptr(b), ptr(bias), None, None, 0, 0, ptr(resid),
                                          ptr(out), None, 0, 0, 0, ptr(stats), None))
    torch.cuda.synchronize()
    x = out.float()
    assert torch.isfinite(stats).all()
    s1, s2 = stats[:, :, 0].sum(1), stats[:, :, 1].sum(1)
    assert (s1 - x.sum(1)).abs().max().item() <= 1e-3 * N ** 0.5 * 4          # f16 rounding of the stored values
    assert ((s2 - (x * x).sum(1)).abs() / (x * x).sum(1)).max().item() <= 1e-3
    w = N // parts                                                             # every slice on its own, too
    assert (stats[:, :, 0] - x.view(M, parts, w).sum(2)).abs().max().item() <= 0.05


@pytest.mark.parametrize("shape", [(300, 128, 128), (700, 1536, 384), (1500, 5120, 1280), (2900, 4096, 1024)])
def test_layernorm_folded_gelu_gemm(ctx, shape):
    """Consuming side (epi 5): f16 stream x f16 gamma-scaled weights, rstd (acc - mean c1) + c2, GELU -- against
    gelu(LayerNorm(x) W^T + b) in f32 torch.  The stream has a non-zero mean and per-row scale, as a residual stream does."""
    from whisper_aries_b200 import _lib
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N + 1)
    x = ((torch.randn(M, K, generator=g) * (0.5 + 2 * torch.rand(M, 1, generator=g)) + 0.8)).cuda().half()
    W = (torch.randn(N, K, generator=g) * 0.04).cuda()
    bias = (torch.randn(N, generator=g) * 0.1).cuda()
    gamma = (1 + 0.1 * torch.randn(K, generator=g)).cuda()
    beta = (0.1 * torch.randn(K, generator=g)).cuda()
    Wf = (W * gamma).half()
    c1 = Wf.float().sum(1)
    c2 = (W.double() @ beta.double()).float() + bias
    parts = ctx.lib.aries_test_gemm_stats_parts(K)
    xs = x.float().view(M, parts, K // parts)
    stats = torch.stack([xs.sum(2), (xs * xs).sum(2)], dim=2).contiguous()
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(ctx.lib.aries_test_gemm_ln(ctx.handle, 5, M, N, K, ptr(x), ptr(Wf), ptr(c2), ptr(c1), ptr(stats), parts, K,
                                          None, ptr(out), None, 0, 0, 0, None, None))
    torch.cuda.synchronize()
    y = torch.nn.functional.layer_norm(x.float(), (K,), gamma, beta, 1e-5)
    ref = torch.nn.functional.gelu(y @ W.t() + bias)
    err = (out.float() - ref).abs().max().item()
    assert err <= 4e-3 + ref.abs().max().item() * 2 ** -7, err                 # bf16 output + f16 weight rounding


def test_layernorm_folded_qkv_split(ctx):
    from whisper_aries_b200 import _lib
    B, T, d, t_pad = 2, 300, 256, 304
    M, N, K = B * T, 3 * d, d
    g = torch.Generator().manual_seed(9)
    x = (torch.randn(M, K, generator=g) * 1.5 - 0.4).cuda().half()
    W = (torch.randn(N, K, generator=g) * 0.05).cuda()
    bias = (torch.randn(N, generator=g) * 0.1).cuda()
    gamma = (1 + 0.1 * torch.randn(K, generator=g)).cuda()
    beta = (0.1 * torch.randn(K, generator=g)).cuda()
    Wf = (W * gamma).half()
    c1, c2 = Wf.float().sum(1), (W.double() @ beta.double()).float() + bias
    parts = ctx.lib.aries_test_gemm_stats_parts(K)
    xs = x.float().view(M, parts, K // parts)
    stats = torch.stack([xs.sum(2), (xs * xs).sum(2)], dim=2).contiguous()
    qk = torch.full((M, 2 * d), float("nan"), device="cuda", dtype=torch.bfloat16)
    vt = torch.zeros((B, d // 64, 64, t_pad), device="cuda", dtype=torch.bfloat16)
    _lib.check(ctx.lib.aries_test_gemm_ln(ctx.handle, 6, M, N, K, ptr(x), ptr(Wf), ptr(c2), ptr(c1), ptr(stats), parts, K,
                                          None, ptr(qk), ptr(vt), 2 * d, T, t_pad, None, None))
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.float(), (K,), gamma, beta, 1e-5) @ W.t() + bias
    tol = 4e-3 + ref.abs().max().item() * 2 ** -7
    assert (qk.float() - ref[:, :2 * d]).abs().max().item() <= tol
    v = ref[:, 2 * d:].view(B, T, d // 64, 64).permute(0, 2, 3, 1)
    assert (vt[..., :T].float() - v).abs().max().item() <= tol


def test_epilogues_write_only_their_rows_and_are_deterministic(ctx):
    """The epilogues store with 256-bit accesses straight from the accumulator rows: rows past M (the last tile is partly
    empty: M = 1000 is not a multiple of 128) must stay untouched in the output AND in the statistics, with guard zones
    on both sides of every buffer, and two launches must give identical bytes."""
    from whisper_aries_b200 import _lib
    M, N, K, G = 1000, 1280, 1280, 64                      # G guard rows before and after
    g = torch.Generator().manual_seed(77)
    a = (torch.randn(M, K, generator=g) * 0.5).cuda().bfloat16()
    b = (torch.randn(N, K, generator=g) * 0.05).cuda().bfloat16()
    bias = torch.randn(N, generator=g).cuda()
    resid_full = torch.full((M + 2 * G, N), 7.0, device="cuda", dtype=torch.float16)
    resid_full[G:G + M] = torch.randn(M, N, generator=g).cuda().half()
    parts = ctx.lib.aries_test_gemm_stats_parts(N)
    runs = []
    for _ in range(2):
        out_full = resid_full.clone()                        # in place on the residual stream, as the encoder runs it
        stats_full = torch.full((M + 2 * G, parts, 2), float("nan"), device="cuda")
        out, stats = out_full[G:G + M], stats_full[G:G + M]
        _lib.check(ctx.lib.aries_test_gemm_ln(ctx.handle, 2, M, N, K, ptr(a), ptr(b), ptr(bias), None, None, 0, 0, ptr(out),
                                              ptr(out), None, 0, 0, 0, ptr(stats), None))
        torch.cuda.synchronize()
        assert (out_full[:G] == 7.0).all() and (out_full[G + M:] == 7.0).all()
        assert torch.isnan(stats_full[:G]).all() and torch.isnan(stats_full[G + M:]).all() and torch.isfinite(stats).all()
        runs.append((out.clone(), stats.clone()))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
    ref = a.float() @ b.float().t() + bias + resid_full[G:G + M].float()
    assert (runs[0][0].float() - ref).abs().max().item() <= 2e-3 + ref.abs().max().item() * 2 ** -10
    # LayerNorm-folded consumer on the same ragged M, fed by those statistics
    W = (torch.randn(N, K, generator=g) * 0.04).cuda()
    Wf = W.half()
    c1, c2 = Wf.float().sum(1).contiguous(), torch.zeros(N, device="cuda")
    x = runs[0][0].contiguous()
    o_full = torch.full((M + 2 * G, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    o = o_full[G:G + M]
    _lib.check(ctx.lib.aries_test_gemm_ln(ctx.handle, 5, M, N, K, ptr(x), ptr(Wf), ptr(c2), ptr(c1), ptr(runs[0][1]), parts, K,
                                          None, ptr(o), None, 0, 0, 0, None, None))
    torch.cuda.synchronize()
    assert torch.isnan(o_full[:G].float()).all() and torch.isnan(o_full[G + M:].float()).all()
    want = torch.nn.functional.gelu(torch.nn.functional.layer_norm(x.float(), (K,), None, None, 1e-5) @ W.t())
    assert (o.float() - want).abs().max().item() <= 4e-3 + want.abs().max().item() * 2 ** -7
