#!/usr/bin/env python
"""Hot-path benchmark: log-mel + Whisper large-v3 encoder over 30-s windows, audio-seconds per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--windows 64] [--model large-v3]
    python bench.py --mode decode [--windows 64] [--tokens 64]      # row f1: greedy generate on the encoder output
    python bench.py --mode transcribe [--windows 64] [--tokens 64]  # int16 PCM -> token ids through the scheduler

A step = one pass of the hot path over one batch of synthetic 30-s / 16 kHz windows per GPU (BASELINE.json config 3:
"large-v3 batch of 64 x 30-s chunks"; weak scaling: every rank owns 64 windows, no collective on the data path).
  value   whole-job audio-s/s with the PCM already resident in HBM (fused PCM -> log-mel -> encoder on the device)
  e2e     the same metric through the public API with HOST buffers: pinned PCM -> H2D -> kernels -> D2H of the bf16
          encoder states, copies inside the timed region (chunk scheduler; micro-batches of --micro-batch windows,
          double-buffered when a shard holds more than one)
  roofline        the dominant kernel (tcgen05 GEMM, all launches of the step) against the measured bf16 peak
  roofline_mel    the log-mel kernel against the measured HBM bandwidth (BASELINE.json: "mel GB/s vs HBM")
  cpu_baseline    the CPU oracle (numpy log-mel + torch fp32 encoder: a port of the reference's algorithm, since
                  faster-whisper / ctranslate2 are not installable offline) on a bounded sample, host cores
Under torchrun (N > 1) every rank runs the same loop on its own GPU; time = max over ranks."""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "audio-sec/sec (log-mel+encoder, large-v3)"
UNIT = "audio-s/s"
WINDOW_SECONDS = 30.0


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def load_traffic():
    """dram bytes per launch from the committed ncu --set full captures (profiles/traffic.json), if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(path)) if os.path.exists(path) else {}


class ClockSampler:
    """Samples SM clock and throttle reasons while the timed region runs (pynvml; nvidia-smi as a fallback)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                     "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
            while not self._stop.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                self._stop.wait(0.05)
        except Exception as exc:                                   # pragma: no cover
            self.reasons.add(f"sampler_error:{type(exc).__name__}")

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thread.join(timeout=2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_reference_step(n_windows: int, model: str, seed0: int = 0, int8: bool = True):
    """The reference algorithm on the host cores: oracle log-mel (numpy) + oracle encoder (torch, all threads) with,
    by default, dynamic-int8 Linear layers -- the nearest stand-in for the reference's faster-whisper
    ``compute_type="int8"`` CPU configuration (ref: final_optimized_transcriber.py:205) -- or plain fp32.
    This is the ONLY place bench.py executes oracle/ code, and only as the thing timed for the CPU baseline."""
    import numpy as np
    import torch
    from oracle import encoder as oenc, logmel as omel, synth as osynth
    shape = osynth.SHAPES[model]
    w = cpu_reference_step.cache.get(model)
    if w is None:
        w = {k: torch.from_numpy(v) for k, v in osynth.encoder_weights(shape, 1234).items()}
        cpu_reference_step.cache[model] = w
    q = None
    if int8:
        q = cpu_reference_step.cache.get(model + "/int8")
        if q is None:
            q = cpu_reference_step.cache[model + "/int8"] = oenc.build_int8_linears(w, shape)
    pcm = osynth.batch_signals(n_windows, seed0)
    t0 = time.perf_counter()
    feats = np.stack([omel.log_mel_window(x, shape.n_mels) for x in pcm])
    t1 = time.perf_counter()
    out = oenc.encoder_forward(feats, w, shape, int8_linears=q)
    t2 = time.perf_counter()
    return {"mel_s": t1 - t0, "enc_s": t2 - t1, "total_s": t2 - t0, "checksum": float(out.abs().mean())}


cpu_reference_step.cache = {}


def reference_backend():
    """faster-whisper / ctranslate2 if some future image has them (also under baseline/_ref); otherwise the port."""
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    try:
        import ctranslate2  # noqa: F401
        import faster_whisper  # noqa: F401
        return "reference"
    except Exception:
        return "port"
    finally:
        sys.path.pop(0)


def run_reference_decode(args, world: int) -> int:
    """--impl reference --mode decode: the oracle decoder (fp32 torch, all host threads) greedy-decoding one window."""
    import torch
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    n = 24
    for _ in range(max(1, args.warmup)):
        cpu_decode_sample(args.model, 4)
    t0 = time.perf_counter()
    toks = 0
    for _ in range(args.steps):
        tps, dt = cpu_decode_sample(args.model, n)
        toks += n
    dt = time.perf_counter() - t0
    value = toks / dt
    sample = (f"1 window x {n} sampled tokens per step: oracle fp32 decoder (torch CPU, whole prefix recomputed per token), "
              f"{args.model} shape, random-init weights; ctranslate2 importable: "
              f"{reference_backend() == 'reference'}")
    line = {"impl": "reference", "metric": f"decoded tokens/sec (greedy generate, {args.model} decoder)", "value": value,
            "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.model} decoder, greedy generate (reference arm: bounded sample of 1 window x {n} "
                                   "tokens per step on the host CPU)", "windows_per_gpu": args.windows,
                       "tokens_per_window": args.tokens},
            "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def run_reference(args, rank: int, world: int) -> int:
    if rank != 0:
        return 0
    if args.mode == "decode":
        return run_reference_decode(args, world)
    import torch
    kind = reference_backend()
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is rank 0 alone and owns the host
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    # (a real faster-whisper leg would need CT2-format weights; with random-init weights only the port can run)
    sample_windows = 2
    for _ in range(args.warmup):
        cpu_reference_step(sample_windows, args.model)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(sample_windows, args.model)
    dt = time.perf_counter() - t0
    value = args.steps * sample_windows * WINDOW_SECONDS / dt
    cores = torch.get_num_threads()
    sample = (f"{sample_windows} synthetic 30-s windows per step (oracle port: numpy log-mel + torch CPU encoder with "
              f"dynamic-int8 Linear layers as the stand-in for faster-whisper compute_type=int8, {args.model} shape, "
              f"random-init weights); faster-whisper/ctranslate2 importable: {kind == 'reference'}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.model} log-mel+encoder, {args.windows} x 30-s windows per GPU per step "
                                   f"(reference arm: bounded sample of {sample_windows} windows per step on the host CPU)",
                       "windows_per_gpu": args.windows, "window_seconds": 30, "sample_rate": 16000},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def cpu_decode_sample(model: str, n_tokens: int):
    """Row f1 CPU baseline: the oracle decoder (torch fp32, all host threads, whole-prefix recompute per token as the
    oracle is written for clarity) greedy-decoding ONE window for a few tokens.  bench.py's only other use of oracle/."""
    import torch
    from oracle import synth as osynth, whisper_decoder as wd
    from whisper_aries_b200 import synthetic
    shape = osynth.DEC_SHAPES[model]
    tok = osynth.WhisperTokens.for_vocab(shape.vocab)
    dec = cpu_decode_sample.cache.get(model)
    if dec is None:                                             # weights are drawn once, outside every timed region
        dec = cpu_decode_sample.cache[model] = wd.Decoder(synthetic.decoder_weights_fast(synthetic.DEC_SHAPES[model], 0), shape)
    enc = torch.randn(1, shape.n_audio_ctx, shape.d_model)
    prompt = [tok.sot, tok.first_lang, tok.transcribe]
    t0 = time.perf_counter()
    res = wd.generate(dec, enc, [prompt], tok, wd.GenerateOptions(max_length=len(prompt) + n_tokens, suppress_tokens=[tok.eot]))
    dt = time.perf_counter() - t0
    return len(res[0]["sequences_ids"]) / dt, dt


cpu_decode_sample.cache = {}


def run_decode(args, rank: int, local_rank: int, world: int, sub: bool = False):
    """--mode decode: greedy ``generate`` (ctranslate2 Whisper.generate, beam 1) for `windows` windows per GPU and
    `tokens` sampled positions per window on a resident bf16 encoder output (synthetic, LayerNorm-ed scale), random-init
    large-v3-shaped decoder.  Weak scaling, no collective.  HBM-bound: per step the decoder reads its weights once and
    every window's cross-attention keys / values (245.8 MB per window for large-v3) once."""
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: whisper_aries_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if sub:
        world = 1                       # a sub-object of the default line: rank 0's GPU only, no process group touched
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from whisper_aries_b200 import WhisperDecoder, synthetic
    shape = synthetic.DEC_SHAPES[args.model]
    tok = synthetic.WhisperTokens.for_vocab(shape.vocab)
    B, T = args.windows, args.tokens
    dec = WhisperDecoder(shape, synthetic.decoder_weights_fast(shape, 0), tokens=tok, device=f"cuda:{local_rank}",
                         max_batch=min(B, 128))
    g = torch.Generator().manual_seed(1000 + rank)
    enc = torch.randn(B, shape.n_audio_ctx, shape.d_model, generator=g).to(dev).bfloat16()
    prompt = [tok.sot, tok.first_lang, tok.transcribe]
    L = len(prompt) + T

    def step():
        # EOT suppressed: every window decodes all T positions, so tokens per step are known
        return dec.generate(enc, [prompt] * B, max_length=L, suppress_tokens=[tok.eot])

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    kv_ms = loop_ms = 0.0
    steps_run = kernels = 0
    with ClockSampler(local_rank) as clocks:
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = step()
            st = dec.last_stats()
            kv_ms += st["cross_kv_ms"]
            loop_ms += st["decode_ms"]
            steps_run += st["steps"]
            kernels += st["steps"] * st["kernels_per_step"] + st["cross_kv_kernels"]
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
        barrier()
    assert all(len(r.sequences_ids[0]) == T for r in res)
    t = torch.tensor([kv_ms + loop_ms, wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    peaks = load_peaks()
    traffic = load_traffic()
    # the dominant kernel of the step (cross-attention over every window's cached keys / values: 36-60 % of the step)
    # timed on its own with CUDA events through the library's kernel-level hook, same shape as inside the step
    import ctypes
    from whisper_aries_b200 import _lib
    ctx = _lib.Context.get(local_rank)
    H, A2 = shape.n_heads, shape.n_audio_ctx
    q = torch.randn(B, shape.d_model, device=dev).bfloat16()
    kv = torch.randn(B * A2, 2 * shape.d_model, device=dev).bfloat16()
    xo = torch.empty((B, shape.d_model), device=dev, dtype=torch.bfloat16)
    want = -(-4 * ctx.sm_count // (B * H))
    xs = max(1, min(8, want))

    def xattn():
        _lib.check(ctx.lib.aries_test_decode_attention(ctx.handle, q.data_ptr(), shape.d_model, kv.data_ptr(),
                                                       kv.data_ptr() + shape.d_model * 2, A2, 2 * shape.d_model, None, None, 0,
                                                       None, A2, B, H, xo.data_ptr(), shape.d_model, xs, None))
    for _ in range(3):
        xattn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_rep = 20
    e0.record()
    for _ in range(n_rep):
        xattn()
    e1.record()
    torch.cuda.synchronize(dev)
    xattn_ms = e0.elapsed_time(e1) / n_rep
    xattn_bytes = B * A2 * 2 * shape.d_model * 2
    del q, kv, xo
    tokens_total = world * B * T * args.steps
    d, f, Lr, V, A = shape.d_model, shape.d_ffn, shape.n_layers, shape.vocab, shape.n_audio_ctx
    w_bytes = (14 * d * d * Lr + V * d) * 2
    xkv_bytes = B * Lr * A * 2 * d * 2
    self_bytes = B * Lr * (L / 2) * 2 * d * 2
    ms_per_token_step = loop_ms / max(steps_run, 1)
    gbs = (w_bytes + xkv_bytes + self_bytes) / (ms_per_token_step * 1e-3) / 1e9
    cpu = None
    if not args.no_cpu_baseline:
        torch.set_num_threads(max(1, os.cpu_count() or 1))
        n = 40
        tps, dt = cpu_decode_sample(args.model, n)
        cpu = {"value": tps, "unit": "tokens/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"1 window x {n} tokens in {dt:.1f} s: oracle fp32 decoder (torch CPU, prefix recomputed per token), "
                         f"{args.model} shape, random-init weights; ctranslate2 not installable offline"}
    line = {"metric": f"decoded tokens/sec (greedy generate, {args.model} decoder)", "value": tokens_total / (dev_ms * 1e-3),
            "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.model} decoder, greedy generate for {B} windows per GPU x {T} sampled tokens "
                                   "(row f1; prompt <|sot|><|lang|><|transcribe|>, timestamp rules on, EOT suppressed)",
                       "windows_per_gpu": B, "tokens_per_window": T, "d_model": d, "layers": Lr, "vocab": V,
                       "l2": "weights 1.6 GB + cross-attention cache 245.8 MB per window exceed the 126 MB L2"},
            "clocks": clocks.summary(),
            "e2e": {"value": tokens_total / (wall_ms * 1e-3), "unit": "tokens/s",
                    "h2d_bytes_per_step": int(B * (len(prompt) + shape.n_text_ctx) * 4),
                    "d2h_bytes_per_step": int(B * (shape.n_text_ctx * 4 + 8)),
                    "api": "WhisperDecoder.generate(encoder_output, prompts) wall clock (prompt upload, token download, "
                           "host-side result assembly inside the timed region; encoder output resident, as upstream keeps it)"},
            "gpu_launches": int(kernels),
            "roofline": {"kernel": f"decode_attention_kernel (cross-attention of one layer: {B} windows x {H} heads x {A2} keys; "
                                   f"{shape.n_layers} launches per token step, the largest share of it)",
                         "bound": "hbm", "achieved": xattn_bytes / (xattn_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": xattn_bytes / (xattn_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                         "peak_source": peaks["source"], "bytes_per_launch": xattn_bytes, "ms_per_launch": xattn_ms,
                         "share_of_step": shape.n_layers * xattn_ms / ms_per_token_step,
                         "traffic": traffic.get("decode_attention_kernel") if B == 64 and args.model == "large-v3" else None},
            "roofline_step": {"kernel": f"whole token step ({st['kernels_per_step']} kernels: skinny tcgen05 GEMMs, single-query "
                                        "attention over the caches, LayerNorm, sampling)",
                              "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                              "frac": gbs / peaks["hbm_gbs"], "peak_source": peaks["source"],
                              "bytes_per_step": {"weights": w_bytes, "cross_kv": xkv_bytes, "self_kv_mean": int(self_bytes)},
                              "ms_per_step": ms_per_token_step},
            "phases": {"cross_kv_projection_ms": kv_ms / args.steps, "decode_loop_ms": loop_ms / args.steps,
                       "ms_per_token_step": ms_per_token_step, "kernels_per_token_step": st["kernels_per_step"]},
            "cpu_baseline": cpu}
    if sub:
        del dec, enc
        torch.cuda.empty_cache()
        return line
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_transcribe(args, rank: int, local_rank: int, world: int, sub: bool = False, model=None):
    """--mode transcribe: the whole widened path through the public API -- pinned int16 PCM windows -> ChunkScheduler ->
    gpu_transcribe_worker (H2D, s16 -> f32, log-mel, encoder, cross K|V, greedy decode of `tokens` ids per window) -> token
    rows on the host.  Wall clock, host buffers, copies inside the timed region; audio-seconds per second."""
    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: whisper_aries_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if sub:
        world = 1
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from whisper_aries_b200 import ChunkScheduler, WhisperDecoder, WhisperModel, gpu_transcribe_worker, synthetic
    eshape, dshape = synthetic.SHAPES[args.model], synthetic.DEC_SHAPES[args.model]
    B, T = args.windows, args.tokens
    if model is None:
        w = dict(synthetic.encoder_weights(eshape, 1234))
        w.update(synthetic.decoder_weights_fast(dshape, 0))
        model = WhisperModel(eshape, w, device="cuda", device_index=local_rank, decoder_shape=dshape,
                             max_batch=min(args.micro_batch, 128))
        del w
    elif model.decoder is None:                     # sub-object: reuse the encoder replica already resident on this GPU
        model.decoder = WhisperDecoder(dshape, synthetic.decoder_weights_fast(dshape, 0), device=f"cuda:{local_rank}",
                                       max_batch=min(args.micro_batch, 128))
    tok = model.decoder.tokens
    prompt = [tok.sot, tok.first_lang, tok.transcribe]
    L = len(prompt) + T
    base = synthetic.batch_signals(min(B, 12), first_seed=rank * B)
    pcm = np.concatenate([base] * (-(-B // base.shape[0])))[:B]
    pcm16 = torch.from_numpy(np.clip(np.round(pcm * 32768.0), -32768, 32767).astype(np.int16)).pin_memory()
    out = torch.zeros((B, L + 1), dtype=torch.int32).pin_memory()
    sched = ChunkScheduler([gpu_transcribe_worker(model, prompt, max_length=L, micro_batch=args.micro_batch,
                                                  suppress_tokens=[tok.eot])])

    def step():
        res = sched.run(pcm16, out)
        if not all(r.success for r in res):
            raise RuntimeError(f"transcribe step failed: {[r.error for r in res if not r.success]}")

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    with ClockSampler(local_rank) as clocks:
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        barrier()
    assert int(out[:, 0].min()) == T
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    if rank == 0:
        value = world * B * WINDOW_SECONDS * args.steps / dt
        st = model.decoder.last_stats()
        line = {"metric": f"audio-sec/sec (PCM -> token ids: log-mel + encoder + greedy decode, {args.model})", "value": value,
                "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"{args.model}: {B} x 30-s int16 windows per GPU per step -> {T} token ids each "
                                       f"(rows a + f1 + f3 through ChunkScheduler / gpu_transcribe_worker, micro-batch "
                                       f"{args.micro_batch})", "windows_per_gpu": B, "tokens_per_window": T},
                "clocks": clocks.summary(),
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(pcm16.numel() * 2),
                        "d2h_bytes_per_step": int(out.numel() * 4),
                        "api": "ChunkScheduler(gpu_transcribe_worker(WhisperModel)).run(pinned int16 pcm, pinned int32 tokens)"},
                "gpu_launches": None,
                "tokens_per_s": world * B * T * args.steps / dt,
                "last_micro_batch": {"cross_kv_ms": st["cross_kv_ms"], "decode_ms": st["decode_ms"], "steps": st["steps"],
                                     "kernels_per_step": st["kernels_per_step"]}}
        if sub:
            sched.close()
            model.decoder.close()
            model.decoder = None
            torch.cuda.empty_cache()
            return line
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def build_replicas(model_name: str, weights: dict, n_gpus: int, first=None):
    """One WhisperModel replica per GPU of this process (full weight copy each: 1.27 GB bf16 for large-v3), built in
    parallel threads -- the ctypes call that converts and uploads the weights releases the GIL."""
    from whisper_aries_b200 import WhisperModel
    models = [first] + [None] * (n_gpus - 1) if first is not None else [None] * n_gpus
    errs = []

    def make(i):
        try:
            models[i] = WhisperModel(model_name, weights, device="cuda", device_index=i)
        except Exception as exc:                                   # pragma: no cover
            errs.append(f"gpu {i}: {type(exc).__name__}: {exc}")

    ts = [threading.Thread(target=make, args=(i,)) for i in range(n_gpus) if models[i] is None]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errs:
        raise RuntimeError("; ".join(errs))
    return models


def measure_inprocess_job(models, n_windows: int, seed0: int, reps: int = 3, micro_batch: int | None = None) -> dict:
    """north_star (3) / BASELINE configs 4 and 5: ONE process shards ``n_windows`` 30-s windows over ``len(models)`` GPUs
    (ChunkScheduler, one worker thread + replica per GPU, static block partition, no collective) and gathers the bf16
    encoder states into ONE pinned host buffer.  Strong scaling: the job is fixed, wall clock runs from the first H2D
    to the last row on the host.  Returns the best and median of ``reps`` runs after one warm-up run."""
    import numpy as np
    import torch
    from whisper_aries_b200 import synthetic
    from whisper_aries_b200.scheduler import ChunkScheduler, gpu_worker, partition_windows
    shape = models[0].shape
    n_gpus = len(models)
    base = synthetic.batch_signals(min(n_windows, 12), first_seed=seed0)
    pcm = torch.from_numpy(np.concatenate([base] * (-(-n_windows // base.shape[0])))[:n_windows].copy()).pin_memory()
    out = torch.empty((n_windows, shape.n_ctx, shape.d_model), dtype=torch.bfloat16).pin_memory()
    shard = -(-n_windows // n_gpus)
    mb = micro_batch or min(64, max(8, -(-shard // 2)))           # two micro-batches per shard: copies overlap compute
    times = []
    with ChunkScheduler([gpu_worker(m, micro_batch=mb) for m in models]) as sched:
        for it in range(reps + 1):
            for m in models:
                torch.cuda.synchronize(m.encoder.device)
            t0 = time.perf_counter()
            res = sched.run(pcm, out)
            dt = time.perf_counter() - t0
            if not all(r.success for r in res):
                raise RuntimeError(f"in-process job failed: {[r.error for r in res if not r.success]}")
            if it:
                times.append(dt)
        per_worker = [round(r.processing_time * 1e3, 2) for r in res]
    times.sort()
    audio = n_windows * WINDOW_SECONDS
    checksum = float(out[:: max(1, n_windows // 8)].float().abs().mean())
    return {"windows": n_windows, "n_gpus": n_gpus, "windows_per_gpu": [b - a for a, b in partition_windows(n_windows, n_gpus)],
            "micro_batch": mb, "reps": reps, "wall_ms_best": times[0] * 1e3, "wall_ms_median": times[len(times) // 2] * 1e3,
            "audio_s_per_s": audio / times[len(times) // 2], "audio_s_per_s_best": audio / times[0],
            "last_run_worker_ms": per_worker, "h2d_bytes": int(pcm.numel() * 4), "d2h_bytes": int(out.numel() * 2),
            "checksum_mean_abs": checksum,
            "api": "ChunkScheduler([gpu_worker(replica_i)]).run(pinned pcm, ONE pinned gather buffer); wall clock, first H2D -> last row on host"}


def parity_object(model, model_name: str, weights: dict) -> dict:
    """One window of the benchmarked model against the CPU oracle, outside every timed region (the oracle is the checker)."""
    import numpy as np
    import torch
    from oracle import encoder as oenc, logmel as omel, synth as osynth
    shape = osynth.SHAPES[model_name]
    pcm = osynth.am_chirp(1)
    want_mel = omel.log_mel(pcm, shape.n_mels)
    got_mel = model.feature_extractor(pcm)
    ref = oenc.encoder_forward(omel.pad_or_trim(want_mel)[None], weights, shape)
    got = model.encode_audio(torch.from_numpy(pcm).to(model.encoder.device)[None])
    cmp = oenc.compare(got.cpu(), ref)
    return {"window": "am_chirp(seed 1), 30 s", "mel_max_abs": float(np.abs(got_mel - want_mel).max()), "mel_tolerance": 1e-4,
            "encoder_max_abs": cmp["max_abs"], "encoder_cosine": cmp["cosine"], "encoder_min_row_cosine": cmp["min_row_cosine"],
            "encoder_bounds": {"cosine": 0.999, "min_row_cosine": 0.999, "max_abs": 0.12},
            "ok": bool(np.abs(got_mel - want_mel).max() <= 1e-4 and cmp["cosine"] >= 0.999 and
                       cmp["min_row_cosine"] >= 0.999 and cmp["max_abs"] <= 0.12),
            "oracle": "oracle/logmel.py + oracle/encoder.py, fp32 on the host (parity unpinned vs faster-whisper/CTranslate2: "
                      "not installable offline); full report: profiles/r02/parity.json"}


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--windows", type=int, default=64, help="30-s windows per GPU per step")
    ap.add_argument("--micro-batch", type=int, default=64,
                    help="windows per H2D/compute/D2H micro-batch (e2e leg); measured e2e audio-s/s at 16 / 32 / 64: 12 150 / 12 415 / 12 568")
    ap.add_argument("--model", default="large-v3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the sub-objects measured after the timed region (parity, config4, config5, decode, transcribe)")
    ap.add_argument("--mode", default="encode", choices=["encode", "decode", "transcribe"],
                    help="encode = the headline log-mel + encoder path; decode = row f1, greedy generate on a resident "
                         "encoder output; transcribe = int16 PCM -> token ids through the scheduler (rows a + f1 + f3)")
    ap.add_argument("--tokens", type=int, default=64, help="--mode decode: sampled tokens per window")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's banner ("NCCL version ...", printed on stdout at debug level
        # VERSION, which this image configures) is raised to WARN, where NCCL honours NCCL_DEBUG_FILE, and sent to stderr
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.mode == "decode":
        return run_decode(args, rank, local_rank, world)
    if args.mode == "transcribe":
        return run_transcribe(args, rank, local_rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: whisper_aries_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)      # timing barrier / max only; the data path has no collective

    from whisper_aries_b200 import WhisperModel, _lib, synthetic
    from whisper_aries_b200.scheduler import ChunkScheduler, gpu_worker, partition_windows

    shape = synthetic.SHAPES[args.model]
    B = args.windows
    weights = synthetic.encoder_weights(shape, 1234)
    model = WhisperModel(shape, weights, device="cuda", device_index=local_rank)
    del weights

    # this rank's shard of the job's windows (static contiguous partition; the seeds make every window distinct)
    start, stop = partition_windows(B * world, world)[rank]
    base = synthetic.batch_signals(min(B, 12), first_seed=start)          # 12 distinct generators, cycled to B windows
    pcm_host = torch.from_numpy(np.concatenate([base] * (-(-B // base.shape[0])))[:B].copy()).pin_memory()
    pcm_dev = pcm_host.to(dev)
    out_dev = torch.empty((B, shape.n_ctx, shape.d_model), dtype=torch.bfloat16, device=dev)
    out_host = torch.empty((B, shape.n_ctx, shape.d_model), dtype=torch.bfloat16).pin_memory()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident():
        model.encoder.encode_pcm(model.feature_extractor, pcm_dev, out=out_dev)

    # ---------------------------------------------------------------- resident leg (value) with per-kernel events
    for _ in range(args.warmup):
        step_resident()
    barrier()
    model.encoder.set_profiling(True)
    model.encoder.collect_profile()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            step_resident()
        ev1.record()
        barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    prof = model.encoder.collect_profile()
    model.encoder.set_profiling(False)
    launches_per_step = model.feature_extractor.last_launches + model.encoder.last_launches
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * B * WINDOW_SECONDS * args.steps / (elapsed_ms * 1e-3)

    # ---------------------------------------------------------------- e2e leg: host buffers through the scheduler
    e2e = None
    if not args.no_e2e:
        sched = ChunkScheduler([gpu_worker(model, micro_batch=args.micro_batch)])
        # Two pinned gather buffers: step i+1 is submitted before step i's rows are awaited (ChunkScheduler.submit), so
        # the download of step i overlaps the kernels of step i+1 -- every step still uploads its PCM from pinned host
        # memory and its result is read on the host (Job.result()) inside the timed region.
        out_hosts = [out_host, torch.empty_like(out_host).pin_memory()]

        def run_e2e(n_steps: int) -> None:
            jobs = []
            for i in range(n_steps):
                jobs.append(sched.submit(pcm_host, out_hosts[i & 1]))
                if len(jobs) > 1:
                    res = jobs.pop(0).result()
                    if not all(r.success for r in res):
                        raise RuntimeError(f"e2e step failed: {[r.error for r in res if not r.success]}")
            for j in jobs:
                res = j.result()
                if not all(r.success for r in res):
                    raise RuntimeError(f"e2e step failed: {[r.error for r in res if not r.success]}")

        run_e2e(max(1, min(args.warmup, 2)))
        barrier()
        t0 = time.perf_counter()
        run_e2e(args.steps)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * WINDOW_SECONDS * args.steps / float(t.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(pcm_host.numel() * 4), "d2h_bytes_per_step": int(out_host.numel() * 2),
               "api": "ChunkScheduler(gpu_worker(WhisperModel)).submit(pinned pcm, pinned out) / Job.result(), two steps in flight",
               "micro_batch": args.micro_batch,
               "pipeline": "uploads, kernels and downloads on three streams, double-buffered across micro-batches and "
                           "across steps: the download of step i overlaps the kernels of step i+1; every step's rows are "
                           "on the host when its Job.result() returns, inside the timed region"}

    # Rank 0 goes on to the in-process multi-GPU jobs (configs 4 / 5) and needs every GPU of the box to itself: the other
    # ranks drop their replicas and park on the rendezvous store (a host-side wait: an NCCL barrier would spin on their GPUs).
    store = dist.distributed_c10d._get_default_store() if world > 1 else None
    if rank != 0:
        torch.cuda.empty_cache()                         # this rank's GPU stays idle from here on (its ~5 GB stay allocated)
        try:
            store.wait(["aries_bench_extras_done"], __import__("datetime").timedelta(seconds=1500))
        except Exception:
            pass
        dist.destroy_process_group()
        return 0

    # ---------------------------------------------------------------- roofline objects (rank 0's kernels)
    peaks = load_peaks()
    traffic = load_traffic()
    d, f, T, L = shape.d_model, shape.d_ffn, shape.n_ctx, shape.n_layers
    c_pad = (shape.n_mels + 63) // 64 * 64
    gemm_flops_per_step = B * (2.0 * 3000 * 3 * c_pad * d + 2.0 * T * 3 * d * d
                               + L * (2.0 * T * d * 3 * d + 2.0 * T * d * d + 4.0 * T * d * f))
    gemm_classes = ["conv1_gemm", "conv2_gemm", "qkv_gemm", "oproj_gemm", "fc1_gemm", "fc2_gemm"]
    gemm_ms = sum(prof[k][0] for k in gemm_classes)
    gemm_launches = sum(prof[k][1] for k in gemm_classes)
    gemm_tflops = gemm_flops_per_step * args.steps / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    peak_tf = peaks["bf16_tflops_sustained"]
    roofline = {"kernel": "gemm_bf16_tcgen05 (all GEMM launches of the step: conv stem, QKV, out-proj, fc1, fc2; since round 2 "
                          "these launches also carry the 64 in-layer LayerNorms -- statistics in the residual epilogues, "
                          "normalisation in the QKV / fc1 epilogues -- which were 6.3 ms of separate kernels in BENCH_r01)",
                "bound": "tensor", "achieved": gemm_tflops, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": gemm_tflops / peak_tf, "peak_source": f"{peaks['source']} (sustained bf16 GEMM)",
                "flops_per_launch": gemm_flops_per_step * args.steps / max(gemm_launches, 1),
                "ms_per_launch": gemm_ms / max(gemm_launches, 1), "launches": gemm_launches,
                "traffic": traffic.get("gemm_bf16_tcgen05")}
    # log-mel: the WHOLE mel path of the step (every launch between PCM and the conv1 operand), against the HBM figure
    # BASELINE.json asks to be quoted (algorithmic bytes = f32 PCM in + f32 [n_mels, 3000] out per window, SURVEY.md 8d,
    # although the fused path writes bf16 time-major and never stores f32 mel).  The kernel is FP32-issue-bound, not
    # HBM-bound: the FMA-pipe / issue-slot utilisation from the committed ncu capture sit beside it.
    mel_bytes = B * (480000 * 4 + shape.n_mels * 3000 * 4)
    mel_classes = ["logmel_tiles", "logmel_clamp", "mel_transpose"]
    mel_ms = sum(prof[k][0] for k in mel_classes)
    mel_launches = sum(prof[k][1] for k in mel_classes)
    mel_passes = max(prof["logmel_tiles"][1], 1)
    mel_gbs = mel_bytes * mel_passes / (mel_ms * 1e-3) / 1e9 if mel_ms > 0 else 0.0
    tiles_ms = prof["logmel_tiles"][0] / mel_passes
    mel_ncu = traffic.get("logmel_tiles_kernel_ncu", {})
    roofline_mel = {"kernel": "whole log-mel path: logmel_tiles_kernel<time-major bf16> + logmel_clamp_tm_kernel "
                              "(PCM f32 -> conv1 operand; no f32 mel in HBM, no transpose launch)",
                    "bound": "hbm", "achieved": mel_gbs, "peak": peaks["hbm_gbs"],
                    "unit": "GB/s", "frac": mel_gbs / peaks["hbm_gbs"], "peak_source": peaks["source"],
                    "bytes_per_launch": mel_bytes, "ms_per_step": mel_ms / mel_passes,
                    "launches_per_step": mel_launches / mel_passes,
                    "tiles_kernel_only": {"ms": tiles_ms, "frac": (mel_bytes / (tiles_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if tiles_ms > 0 else 0.0},
                    "fp32_pipe_frac": mel_ncu.get("fma_pipe_frac"), "issue_slot_frac": mel_ncu.get("issue_active_frac"),
                    "binding": "fp32 issue (ncu: issue slots / FMA pipe, profiles/r02/ncu_full_mel-fused.txt), not HBM",
                    "traffic": traffic.get("logmel_tiles_kernel")}
    attn_ms, attn_n = prof["attention"]
    attn_flops = B * 4.0 * T * T * d
    kernels = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in prof.items()}
    kernels["attention"]["tflops"] = attn_flops * attn_n / (attn_ms * 1e-3) / 1e12 if attn_ms > 0 else 0.0
    encoder_tflops = B * shape.flops_per_window * args.steps / (elapsed_ms * 1e-3) / 1e12

    cpu_baseline = None
    if not args.no_cpu_baseline:
        import torch as _t
        _t.set_num_threads(max(1, os.cpu_count() or 1))              # torchrun pins OMP_NUM_THREADS=1 per rank
        n_sample = 2 if args.model == "large-v3" else 4
        cpu_reference_step(1, args.model)                             # warm the weights / thread pool
        r = cpu_reference_step(n_sample, args.model)
        r32 = cpu_reference_step(n_sample, args.model, int8=False)
        # the reference's own worker default is cpu_threads=2 (ref: final_optimized_transcriber.py:177): timed beside
        # the generous all-core figure, on one window
        all_threads = _t.get_num_threads()
        _t.set_num_threads(2)
        r2t = cpu_reference_step(1, args.model)
        _t.set_num_threads(all_threads)
        cpu_baseline = {"value": n_sample * WINDOW_SECONDS / r["total_s"], "unit": UNIT, "cores": all_threads,
                        "kind": "port", "value_2_threads_reference_default": WINDOW_SECONDS / r2t["total_s"],
                        "sample": f"{n_sample} windows of the same workload: numpy log-mel {r['mel_s']:.2f} s + torch CPU "
                                  f"encoder with dynamic-int8 Linear layers {r['enc_s']:.2f} s (oracle port standing in for "
                                  f"faster-whisper compute_type=int8; faster-whisper/ctranslate2 not installable offline)",
                        "fp32_value": n_sample * WINDOW_SECONDS / r32["total_s"]}

    # ---------------------------------------------------------------- outside every timed region: parity + widened jobs
    del pcm_dev, out_dev, out_host, pcm_host
    if not args.no_e2e:
        del out_hosts
    torch.cuda.empty_cache()
    extras = {}
    if not args.no_extras:
        def guarded(name, fn):
            t0 = time.perf_counter()
            try:
                extras[name] = fn()
            except Exception as exc:                              # an extra must never cost the headline line
                extras[name] = {"error": f"{type(exc).__name__}: {exc}"}
            if isinstance(extras[name], dict):
                extras[name]["bench_seconds"] = round(time.perf_counter() - t0, 2)

        n_dev = min(world, torch.cuda.device_count())
        torch.set_num_threads(max(1, os.cpu_count() or 1))       # the oracle (parity checker) may use the whole host
        weights = synthetic.encoder_weights(shape, 1234)
        guarded("parity", lambda: parity_object(model, args.model, weights))

        def config4():
            # BASELINE config 4: 1-hour stream = 120 windows, strong scaling over the N GPUs of this run (seed 7)
            models = build_replicas(args.model, weights, n_dev, first=model)
            r = measure_inprocess_job(models, 120, 7000)
            if n_dev > 1:
                r["one_gpu"] = {k: v for k, v in measure_inprocess_job(models[:1], 120, 7000).items()
                                if k in ("wall_ms_median", "audio_s_per_s", "micro_batch")}
            r["config"] = f"BASELINE config 4: {args.model}, 120 x 30-s windows (1 h) over {n_dev} GPU(s), one process"
            for m in models[1:]:
                m.encoder.close()
            return r

        guarded("config4", config4)
        del weights

        def config5():
            # BASELINE config 5: Whisper medium (80 bins, d 1024, 16 heads, 24 layers), 128 windows over the N GPUs
            mshape = synthetic.SHAPES["medium"]
            mw = synthetic.encoder_weights(mshape, 1234)
            models = build_replicas("medium", mw, n_dev)
            r = measure_inprocess_job(models, 128, 500)
            r["config"] = f"BASELINE config 5: medium, 128 x 30-s windows over {n_dev} GPU(s), one process (alt-shape path)"
            r["encoder_tflops"] = 128 * mshape.flops_per_window / (r["wall_ms_median"] * 1e-3) / 1e12
            for m in models:
                m.encoder.close()
            return r

        guarded("config5", config5)
        if world == 1:
            import copy
            sub = copy.copy(args)
            sub.steps, sub.warmup, sub.tokens, sub.no_cpu_baseline, sub.windows, sub.micro_batch = 2, 1, 16, True, 64, 64
            guarded("decode", lambda: run_decode(sub, 0, local_rank, 1, sub=True))
            guarded("transcribe", lambda: run_transcribe(sub, 0, local_rank, 1, sub=True, model=model))
    if store is not None:
        store.set("aries_bench_extras_done", "1")

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.model} log-mel+encoder, {B} x 30-s windows per GPU per step (BASELINE.json config 3)",
                       "windows_per_gpu": B, "window_seconds": 30, "sample_rate": 16000, "n_mels": shape.n_mels,
                       "d_model": d, "layers": L, "parallelism": f"data-parallel over windows, {world} rank(s), no collective",
                       "l2": "per-step working set (PCM 123 MB + ~3.3 GB of activations) exceeds the 126 MB L2; no flush needed"},
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
            "roofline": roofline, "roofline_mel": roofline_mel, "encoder_tflops": encoder_tflops,
            "encoder_frac_of_bf16_peak": encoder_tflops / peak_tf, "kernels": kernels, "cpu_baseline": cpu_baseline}
    line.update(extras)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
