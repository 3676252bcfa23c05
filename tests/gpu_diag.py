"""Per-kernel diagnostics on a real B200 (run under gpurun; not collected by pytest).

    python tests/gpu_diag.py <what> [...]     what in: mel gemm ln attn enc-micro enc-tiny enc-large melperf gemmperf

Each section prints error statistics against the oracle (test infrastructure) / a torch fp32 reference and, for the
*perf sections, CUDA-event timings.  Sections are independent so a failure in one does not hide the others."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import encoder as oenc, logmel as omel, synth as osynth   # noqa: E402  (checker only)
from whisper_aries_b200 import FeatureExtractor, WhisperEncoder, _lib, synthetic   # noqa: E402

dev = torch.device("cuda:0")


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    return min(ts), sum(ts) / len(ts)


def diag_mel():
    for n_mels in (80, 128):
        fe = FeatureExtractor(feature_size=n_mels)
        for name, x in (("tone", osynth.tone_noise(0)), ("chirp", osynth.am_chirp(1)), ("gapped", osynth.gapped(2)),
                        ("short1234", osynth.window_signal(7, 1234)), ("s8000", osynth.window_signal(7, 8000)),
                        ("s29600", osynth.window_signal(7, 29600)), ("zeros", np.zeros(480000, np.float32))):
            ref = omel.log_mel(x, n_mels)
            got = fe(x)
            err = np.abs(got - ref)
            print(f"mel n_mels={n_mels} {name:10s} shape={got.shape} max_err={err.max():.3e} "
                  f"mean_err={err.mean():.3e} argmax={np.unravel_index(err.argmax(), err.shape)}", flush=True)
        # batch + device path + frames_out=3000
        xs = osynth.batch_signals(5, 10)
        got = fe(torch.from_numpy(xs).to(dev), frames_out=3000).cpu().numpy()
        ref = np.stack([omel.log_mel_window(x, n_mels) for x in xs])
        print(f"mel n_mels={n_mels} batch5 device frames_out=3000 max_err={np.abs(got - ref).max():.3e}", flush=True)


def diag_melperf():
    fe = FeatureExtractor(feature_size=128)
    for B in (1, 16, 64, 148):
        xs = torch.from_numpy(osynth.batch_signals(min(B, 6), 0)).to(dev)
        xs = xs.repeat((B + xs.shape[0] - 1) // xs.shape[0], 1)[:B].contiguous()
        out = torch.empty((B, 128, 3000), device=dev)
        lib, h = fe._ctx.lib if fe._handle else None, None
        fe(xs[:1])
        lib, h = fe._ctx.lib, fe._handle

        def run():
            _lib.check(lib.aries_logmel_run(h, xs.data_ptr(), B, 480000, xs.stride(0), 160, out.data_ptr(), 3000, None))
        best, avg = timed(run, iters=20)
        gb = B * 3.456e6 / 1e9
        print(f"melperf B={B} best={best * 1e3:.1f}us avg={avg * 1e3:.1f}us  {gb / (best * 1e-3):.0f} GB/s best "
              f"{gb / (avg * 1e-3):.0f} GB/s avg  ({best * 1e3 / B:.2f} us/window)", flush=True)


def diag_melfused():
    """The fused PCM -> conv1-operand kernels of aries_encode_pcm (per-kernel CUDA events of the library's profiler),
    next to the f32 FeatureExtractor path, same process and clocks: 64 windows, 128 mel bins."""
    from whisper_aries_b200 import WhisperModel, synthetic
    B = 64
    shape = synthetic.EncoderShape("mel128-toy", 128, 128, 2, 2, 512)
    model = WhisperModel(shape, synthetic.encoder_weights(shape, 1), device="cuda", device_index=0)
    xs = torch.from_numpy(osynth.batch_signals(6, 0)).to(dev).repeat(11, 1)[:B].contiguous()
    fe = model.feature_extractor
    out = torch.empty((B, 128, 3000), device=dev)
    fe(xs[:1])
    lib, h = fe._ctx.lib, fe._handle
    for rep in range(3):
        for _ in range(3):
            model.encode_audio(xs)
        model.encoder.set_profiling(True)
        model.encoder.collect_profile()
        n = 20
        for _ in range(n):
            model.encode_audio(xs)
        torch.cuda.synchronize()
        prof = model.encoder.collect_profile()
        model.encoder.set_profiling(False)

        def run():
            _lib.check(lib.aries_logmel_run(h, xs.data_ptr(), B, 480000, xs.stride(0), 160, out.data_ptr(), 3000, None))
        best, avg = timed(run, iters=20)
        print(f"melfused B={B}: fused tiles {prof['logmel_tiles'][0] / n * 1e3:.1f} us + clamp {prof['logmel_clamp'][0] / n * 1e3:.1f} us"
              f" | f32 path (tiles + clamp) best {best * 1e3:.1f} us avg {avg * 1e3:.1f} us", flush=True)


def gemm_ref(a, b, bias, epi, resid=None, pos=None, pos_rows=0):
    acc = a.float() @ b.float().t() + bias
    if epi in (1, 3):
        acc = torch.nn.functional.gelu(acc)
    if epi == 2:
        acc = acc + resid
    if epi == 3:
        idx = torch.arange(a.shape[0], device=a.device) % pos_rows
        acc = acc + pos[idx]
    return acc


def diag_gemm():
    ctx = _lib.Context.get(0)
    lib = ctx.lib
    g = torch.Generator(device="cpu").manual_seed(0)
    for (M, N, K) in ((128, 128, 64), (128, 128, 256), (128, 256, 128), (300, 384, 128), (1500, 1280, 1280),
                      (3000, 3840, 1280), (1000, 1280, 5120)):
        a = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
        b = (torch.randn(N, K, generator=g) * 0.05).to(dev).bfloat16()
        bias = torch.randn(N, generator=g).to(dev)
        resid = torch.randn(M, N, generator=g).to(dev).half()
        pos_rows = 100
        pos = torch.randn(pos_rows, N, generator=g).to(dev)
        for epi in (0, 1, 2, 3):
            out = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16 if epi < 2 else torch.float16)
            Mx = (M // pos_rows) * pos_rows if epi == 3 else M
            rc = lib.aries_test_gemm(ctx.handle, epi, Mx, N, K, ptr(a), ptr(b), ptr(bias), ptr(resid), ptr(pos), pos_rows,
                                     ptr(out), None, 0, 0, 0, None)
            torch.cuda.synchronize()
            if rc:
                print(f"gemm M={M} N={N} K={K} epi={epi} rc={rc} {_lib.last_error()}", flush=True)
                continue
            ref = gemm_ref(a[:Mx], b, bias, epi, resid[:Mx].float(), pos, pos_rows)
            err = (out[:Mx].float() - ref).abs()
            bad = torch.isnan(out[:Mx].float()).sum().item()
            print(f"gemm M={Mx} N={N} K={K} epi={epi} max_err={err.nan_to_num(1e9).max().item():.3e} "
                  f"ref_max={ref.abs().max().item():.2f} nan={bad}", flush=True)
            if err.nan_to_num(1e9).max().item() > 0.1 and M <= 300:
                e = err.nan_to_num(1e9)
                rows = (e.max(dim=1).values > 0.1).nonzero().flatten()[:16].tolist()
                cols = (e.max(dim=0).values > 0.1).nonzero().flatten()[:16].tolist()
                print("   bad rows", rows, "bad cols", cols, flush=True)
    # QKV split epilogue
    B_, T_, d = 2, 300, 256
    M, N, K = B_ * T_, 3 * d, 128
    a = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
    b = (torch.randn(N, K, generator=g) * 0.05).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    t_pad = 304
    out = torch.zeros((M, 2 * d), device=dev, dtype=torch.bfloat16)
    out2 = torch.zeros((B_, d // 64, 64, t_pad), device=dev, dtype=torch.bfloat16)
    rc = lib.aries_test_gemm(ctx.handle, 4, M, N, K, ptr(a), ptr(b), ptr(bias), None, None, 0, ptr(out), ptr(out2),
                             2 * d, T_, t_pad, None)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t() + bias
    e1 = (out.float() - ref[:, : 2 * d]).abs().max().item()
    v = ref[:, 2 * d:].reshape(B_, T_, d // 64, 64).permute(0, 2, 3, 1)
    e2 = (out2[..., :T_].float() - v).abs().max().item()
    print(f"gemm qkv-split rc={rc} qk_err={e1:.3e} vt_err={e2:.3e} pad_untouched={out2[..., T_:].abs().max().item() == 0}",
          flush=True)


def diag_gemmperf():
    ctx = _lib.Context.get(0)
    lib = ctx.lib
    M = 96000
    for (N, K, epi, name) in ((3840, 1280, 0, "qkv"), (1280, 1280, 2, "o"), (5120, 1280, 1, "fc1"), (1280, 5120, 2, "fc2")):
        a = torch.randn(M, K, device=dev).bfloat16()
        b = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        bias = torch.randn(N, device=dev)
        resid = torch.randn(M, N, device=dev).half() if epi == 2 else None
        out = torch.empty((M, N), device=dev, dtype=torch.bfloat16 if epi < 2 else torch.float16)

        def run():
            lib.aries_test_gemm(ctx.handle, epi, M, N, K, ptr(a), ptr(b), ptr(bias), ptr(resid), None, 0, ptr(out), None,
                                0, 0, 0, None)
        best, avg = timed(run, iters=8)
        fl = 2.0 * M * N * K
        print(f"gemmperf {name} M={M} N={N} K={K} best={best:.3f}ms avg={avg:.3f}ms {fl / best / 1e9:.0f} TFLOP/s best "
              f"{fl / avg / 1e9:.0f} avg", flush=True)
        def run_t():
            torch.matmul(a, b.t())
        best, avg = timed(run_t, iters=8)
        print(f"   torch.matmul (cuBLAS) best={best:.3f}ms {fl / best / 1e9:.0f} TFLOP/s", flush=True)


def diag_ln():
    ctx = _lib.Context.get(0)
    for d in (128, 384, 1024, 1280):
        x = (torch.randn(1000, d, device=dev) * 3 + 0.5).half()
        gm, bt = torch.randn(d, device=dev), torch.randn(d, device=dev)
        y = torch.empty((1000, d), device=dev, dtype=torch.bfloat16)
        rc = ctx.lib.aries_test_layernorm(ctx.handle, ptr(x), ptr(gm), ptr(bt), ptr(y), 1000, d, None)
        torch.cuda.synchronize()
        ref = torch.nn.functional.layer_norm(x.float(), (d,), gm, bt, 1e-5)
        print(f"ln d={d} rc={rc} max_err={(y.float() - ref).abs().max().item():.3e} (bf16 out, ref_max={ref.abs().max().item():.1f})",
              flush=True)


def diag_attn():
    ctx = _lib.Context.get(0)
    g = torch.Generator(device="cpu").manual_seed(1)
    for (B_, T_, H, qs) in ((1, 128, 1, 1.0), (1, 256, 2, 1.0), (1, 300, 3, 1.0), (1, 77, 1, 3.0), (2, 1500, 2, 1.0),
                            (2, 1500, 2, 5.0), (1, 1500, 20, 1.0), (8, 1500, 20, 1.5)):
        d = 64 * H
        t_pad = (T_ + 7) // 8 * 8
        q = torch.randn(B_, T_, H, 64, generator=g).to(dev) * qs
        k = torch.randn(B_, T_, H, 64, generator=g).to(dev)
        v = torch.randn(B_, T_, H, 64, generator=g).to(dev)
        qk = torch.cat([q.reshape(B_ * T_, d), k.reshape(B_ * T_, d)], dim=1).bfloat16().contiguous()
        vt = torch.full((B_, H, 64, t_pad), float("nan"), device=dev, dtype=torch.bfloat16)
        vt[..., :T_] = v.permute(0, 2, 3, 1).bfloat16()
        out = torch.full((B_ * T_, d), float("nan"), device=dev, dtype=torch.bfloat16)
        rc = ctx.lib.aries_test_attention(ctx.handle, ptr(qk), ptr(vt), B_, T_, H, t_pad, ptr(out), None)
        torch.cuda.synchronize()
        qf, kf, vf = (t.bfloat16().float().permute(0, 2, 1, 3) for t in (q, k, v))
        att = torch.softmax(qf @ kf.transpose(-1, -2) / 8.0, dim=-1)
        ref = (att @ vf).permute(0, 2, 1, 3).reshape(B_ * T_, d)
        err = (out.float() - ref).abs()
        print(f"attn B={B_} T={T_} H={H} qscale={qs} rc={rc} max_err={err.nan_to_num(1e9).max().item():.3e} "
              f"nan={torch.isnan(out.float()).sum().item()} ref_max={ref.abs().max().item():.2f}", flush=True)
        if B_ * H * T_ >= 30000:
            def run():
                ctx.lib.aries_test_attention(ctx.handle, ptr(qk), ptr(vt), B_, T_, H, t_pad, ptr(out), None)
            best, avg = timed(run)
            fl = 4.0 * B_ * H * T_ * T_ * 64
            print(f"   attn time best={best * 1e3:.1f}us {fl / best / 1e9:.0f} TFLOP/s", flush=True)


def diag_enc(shape_name, batch=2, check=True):
    shape = synthetic.SHAPES[shape_name]
    w = synthetic.encoder_weights(shape, 1234)
    t0 = time.time()
    enc = WhisperEncoder(shape, w)
    print(f"enc {shape_name}: create {time.time() - t0:.1f}s ws(batch={batch})={enc.workspace_bytes(batch) / 1e6:.0f} MB", flush=True)
    feats = np.stack([omel.log_mel_window(osynth.window_signal(s), shape.n_mels) for s in range(batch)])
    out = enc.encode(feats)
    torch.cuda.synchronize()
    print(f"enc {shape_name}: launches={enc.last_launches} out finite={torch.isfinite(out.float()).all().item()}", flush=True)
    if check:
        ref, layers = oenc.encoder_forward(feats, w, shape, return_layers=True)
        print(f"enc {shape_name}: vs fp32 oracle {oenc.compare(out.cpu(), ref)}", flush=True)
        ref16 = oenc.encoder_forward(feats, w, shape, round_weights_bf16=True)
        print(f"enc {shape_name}: vs oracle with bf16-rounded weights {oenc.compare(out.cpu(), ref16)}", flush=True)
    x = torch.from_numpy(feats).to(dev)

    def run():
        enc.encode(x)
    best, avg = timed(run, iters=5, warm=2)
    fl = batch * shape.flops_per_window
    print(f"enc {shape_name}: batch={batch} best={best:.3f}ms avg={avg:.3f}ms -> {fl / best / 1e9:.1f} TFLOP/s, "
          f"{batch * 30 / (best * 1e-3):.0f} audio-s/s", flush=True)


if __name__ == "__main__":
    torch.cuda.init()
    print(torch.cuda.get_device_name(0), flush=True)
    for what in sys.argv[1:]:
        t0 = time.time()
        print(f"==== {what}", flush=True)
        if what == "mel":
            diag_mel()
        elif what == "melperf":
            diag_melperf()
        elif what == "gemm":
            diag_gemm()
        elif what == "gemmperf":
            diag_gemmperf()
        elif what == "ln":
            diag_ln()
        elif what == "attn":
            diag_attn()
        elif what == "melfused":
            diag_melfused()
        elif what.startswith("enc-"):
            parts = what.split("-", 1)[1].split(":")
            name = parts[0]
            batch = int(parts[1]) if len(parts) > 1 else 2
            diag_enc(name, batch, check=(name != "large-v3" or batch <= 2))
        print(f"==== {what} done in {time.time() - t0:.1f}s", flush=True)
