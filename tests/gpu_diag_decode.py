"""GPU diagnostics for row f1 (run under gpurun): numeric error of the decoder logits against the oracle at the 'mini'
shape, and decode-step timing at the large-v3 shape for several batch sizes / launch modes.

    python tests/gpu_diag_decode.py err
    python tests/gpu_diag_decode.py perf [batch ...]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def fast_decoder_weights(shape, seed=0):
    from whisper_aries_b200 import synthetic
    return synthetic.decoder_weights_fast(shape, seed)


def err():
    from oracle import synth as osynth, whisper_decoder as wd
    from whisper_aries_b200 import WhisperDecoder, synthetic
    for name, seed in (("micro", 32), ("mini", 20)):
        shape = synthetic.DEC_SHAPES[name]
        tok = synthetic.WhisperTokens.for_vocab(shape.vocab)
        batch = 2
        g = torch.Generator().manual_seed(seed + batch)
        enc = torch.randn(batch, shape.n_audio_ctx, shape.d_model, generator=g).bfloat16()
        dec = WhisperDecoder(shape, synthetic.decoder_weights(shape, seed), tokens=tok, max_batch=4)
        oshape = osynth.DEC_SHAPES[name]
        oracle = wd.Decoder(osynth.decoder_weights(oshape, seed), oshape, round_weights_bf16=True)
        prompt = [tok.sot, tok.first_lang + 1, tok.transcribe]
        L = len(prompt) + 12
        ref = wd.generate(oracle, enc.float(), [prompt] * batch, tok, wd.GenerateOptions(max_length=L))
        forced = [r["sequences_ids"] for r in ref]
        res, extras = dec.generate(enc.cuda(), [prompt] * batch, max_length=L, suppress_tokens=[], _forced=forced,
                                   _want_logits=True, return_no_speech_prob=True)
        seqs = torch.tensor([prompt + f for f in forced])[:, :L - 1]
        ref_logits = oracle.logits(seqs, enc.float())
        for t in range(L - 1):
            a, r = torch.from_numpy(extras[0]["logits"][t]), ref_logits[:, t]
            cos = torch.nn.functional.cosine_similarity(a, r, dim=-1).min().item()
            print(f"{name} step {t:2d}: logits cosine {cos:.6f} max-abs err {(a - r).abs().max().item():.4f} "
                  f"scale {r.abs().max().item():.2f} rms err {(a - r).pow(2).mean().sqrt().item():.4f}")
        print(name, "no_speech_prob", [r.no_speech_prob for r in res], [r["no_speech_prob"] for r in ref])


def perf(batches, modes=("graph+pdl", "graph", "stream+pdl", "stream"), stacks=("0",)):
    from whisper_aries_b200 import WhisperDecoder, synthetic
    shape = synthetic.DEC_SHAPES["large-v3"]
    tok = synthetic.WhisperTokens.for_vocab(shape.vocab)
    t0 = time.time()
    w = fast_decoder_weights(shape)
    print(f"weights drawn in {time.time() - t0:.1f} s", flush=True)
    t0 = time.time()
    dec = WhisperDecoder(shape, w, tokens=tok, max_batch=max(batches))
    print(f"decoder created in {time.time() - t0:.1f} s", flush=True)
    del w
    prompt = [tok.sot, tok.first_lang, tok.transcribe]
    wbytes = (14 * shape.d_model ** 2 * shape.n_layers + shape.vocab * shape.d_model) * 2
    for B in batches:
        enc = torch.randn(B, shape.n_audio_ctx, shape.d_model, device="cuda").bfloat16()
        for mode, stack in [(m, s) for s in stacks for m in modes]:
            os.environ["ARIES_DECODE_STACK"] = stack
            os.environ["ARIES_DECODE_GRAPH"] = "1" if mode.startswith("graph") else "0"
            os.environ["ARIES_DECODE_PDL"] = "1" if mode.endswith("pdl") else "0"
            L = 64 + len(prompt)
            for rep in range(2):
                dec.generate(enc, [prompt] * B, max_length=L, suppress_tokens=[tok.eot])     # never stops early
                st = dec.last_stats()
            ms = st["decode_ms"] / st["steps"]
            kv = B * shape.n_layers * shape.n_audio_ctx * 2 * shape.d_model * 2
            self_kv = B * shape.n_layers * (L / 2) * 2 * shape.d_model * 2
            gbs = (wbytes + kv + self_kv) / ms / 1e6
            print(f"B={B:3d} {mode:11s} stack={stack}: {ms:7.3f} ms/step ({st['kernels_per_step']} kernels, {ms * 1e3 / st['kernels_per_step']:.2f} us each) "
                  f"{B / ms * 1e3:9.0f} tok/s  {gbs:7.0f} GB/s of {wbytes / 1e9:.2f}+{kv / 1e9:.2f} GB; cross-KV projection "
                  f"{st['cross_kv_ms']:.2f} ms", flush=True)


def prof(batch, steps=6):
    """Target for ncu: a few decode steps at the large-v3 layer shape (4 layers: same kernels, 8x less to replay),
    launched directly on the stream (no graph, no programmatic launch) so that every kernel is a separate record."""
    from whisper_aries_b200 import WhisperDecoder, synthetic
    os.environ["ARIES_DECODE_GRAPH"] = "0"
    os.environ["ARIES_DECODE_PDL"] = "0"
    shape = synthetic.DecoderShape("large-v3-4l", 51866, 1280, 20, 4, 5120)
    tok = synthetic.WhisperTokens.for_vocab(shape.vocab)
    dec = WhisperDecoder(shape, fast_decoder_weights(shape), tokens=tok, max_batch=batch)
    enc = torch.randn(batch, shape.n_audio_ctx, shape.d_model, device="cuda").bfloat16()
    prompt = [tok.sot, tok.first_lang, tok.transcribe]
    dec.generate(enc, [prompt] * batch, max_length=len(prompt) + steps, suppress_tokens=[tok.eot])
    print("ok", dec.last_stats())


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "err"
    if what == "err":
        err()
    elif what == "pdlmask":
        # which kernel classes should take a programmatic edge at a given batch (ARIES_DECODE_PDL_MASK)
        from whisper_aries_b200 import WhisperDecoder, synthetic
        shape = synthetic.DEC_SHAPES["large-v3"]
        tok = synthetic.WhisperTokens.for_vocab(shape.vocab)
        batches = [int(a) for a in sys.argv[2:]] or [64]
        dec = WhisperDecoder(shape, fast_decoder_weights(shape), tokens=tok, max_batch=max(batches))
        prompt = [tok.sot, tok.first_lang, tok.transcribe]
        os.environ["ARIES_DECODE_GRAPH"] = "1"
        os.environ["ARIES_DECODE_PDL"] = "1"
        for B in batches:
            enc = torch.randn(B, shape.n_audio_ctx, shape.d_model, device="cuda").bfloat16()
            for mask in range(8):
                os.environ["ARIES_DECODE_PDL_MASK"] = str(mask)
                for _ in range(2):
                    dec.generate(enc, [prompt] * B, max_length=64 + len(prompt), suppress_tokens=[tok.eot])
                    st = dec.last_stats()
                print(f"B={B} pdl mask {mask} (1 gemm, 2 attention, 4 rest): {st['decode_ms'] / st['steps']:.3f} ms/step", flush=True)
    elif what == "prof":
        prof(int(sys.argv[2]) if len(sys.argv) > 2 else 64)
    elif what == "stack":
        # persistent stack kernel (<= 8 windows) against the launch-per-op step, graph replay with programmatic launch
        perf([int(a) for a in sys.argv[2:]] or [1, 2, 4, 8], modes=("graph+pdl",), stacks=("1", "0"))
    else:
        perf([int(a) for a in sys.argv[2:]] or [1, 8, 64])
