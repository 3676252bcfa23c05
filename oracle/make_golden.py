"""Regenerates tests/golden/*.npz (oracle; test infrastructure).  Run from the repo root:

    python -m oracle.make_golden

The true reference (faster-whisper 1.1.1 / ctranslate2 4.6.0) is not importable in this sandbox, so
the fixtures hold (a) the oracle's own outputs on the SURVEY section 8d signals — a regression pin —
and (b) the outputs of HF transformers 5.5's independent WhisperFeatureExtractor / WhisperEncoder on
the same inputs, generated HERE by importing transformers (it does not travel to the GPU box as a
test dependency: tests only read the stored arrays).  HF differs from faster-whisper by design in the
last two frames of a 30-s window (pad-to-30-s-then-reflect vs zero-pad-160-then-reflect), so the HF
cross-pin covers frames 0..2997 only.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

from . import encoder, logmel, synth
from .decoder import GreedyProbe

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
FRAME_PICK = np.concatenate([np.arange(0, 2997, 61), np.arange(2994, 3001)])   # includes the edge frames


def _sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def hf_encoder(shape: synth.EncoderShape, w: dict):
    from transformers import WhisperConfig
    from transformers.models.whisper.modeling_whisper import WhisperEncoder
    cfg = WhisperConfig(num_mel_bins=shape.n_mels, d_model=shape.d_model, encoder_layers=shape.n_layers,
                        encoder_attention_heads=shape.n_heads, encoder_ffn_dim=shape.d_ffn,
                        max_source_positions=shape.n_ctx)
    enc = WhisperEncoder(cfg).eval()
    sd = enc.state_dict()
    d = shape.d_model

    def T(k):
        return torch.from_numpy(w[k])

    sd["conv1.weight"], sd["conv1.bias"] = T("encoder/conv1/weight"), T("encoder/conv1/bias")
    sd["conv2.weight"], sd["conv2.bias"] = T("encoder/conv2/weight"), T("encoder/conv2/bias")
    sd["embed_positions.weight"] = T("encoder/position_encodings/encodings")
    for i in range(shape.n_layers):
        p, q = f"encoder/layer_{i}", f"layers.{i}"
        W, b = T(f"{p}/self_attention/linear_0/weight"), T(f"{p}/self_attention/linear_0/bias")
        sd[f"{q}.self_attn.q_proj.weight"], sd[f"{q}.self_attn.q_proj.bias"] = W[:d], b[:d]
        sd[f"{q}.self_attn.k_proj.weight"] = W[d:2 * d]
        sd[f"{q}.self_attn.v_proj.weight"], sd[f"{q}.self_attn.v_proj.bias"] = W[2 * d:], b[2 * d:]
        sd[f"{q}.self_attn.out_proj.weight"] = T(f"{p}/self_attention/linear_1/weight")
        sd[f"{q}.self_attn.out_proj.bias"] = T(f"{p}/self_attention/linear_1/bias")
        sd[f"{q}.self_attn_layer_norm.weight"] = T(f"{p}/self_attention/layer_norm/gamma")
        sd[f"{q}.self_attn_layer_norm.bias"] = T(f"{p}/self_attention/layer_norm/beta")
        sd[f"{q}.final_layer_norm.weight"] = T(f"{p}/ffn/layer_norm/gamma")
        sd[f"{q}.final_layer_norm.bias"] = T(f"{p}/ffn/layer_norm/beta")
        sd[f"{q}.fc1.weight"], sd[f"{q}.fc1.bias"] = T(f"{p}/ffn/linear_0/weight"), T(f"{p}/ffn/linear_0/bias")
        sd[f"{q}.fc2.weight"], sd[f"{q}.fc2.bias"] = T(f"{p}/ffn/linear_1/weight"), T(f"{p}/ffn/linear_1/bias")
    sd["layer_norm.weight"], sd["layer_norm.bias"] = T("encoder/layer_norm/gamma"), T("encoder/layer_norm/beta")
    enc.load_state_dict(sd)
    return enc


def hf_decoder(shape: synth.DecoderShape, w: dict):
    """HF transformers WhisperDecoder carrying the oracle's weights (CT2 variable names -> HF parameter names)."""
    from transformers import WhisperConfig
    from transformers.models.whisper.modeling_whisper import WhisperDecoder
    cfg = WhisperConfig(vocab_size=shape.vocab, d_model=shape.d_model, decoder_layers=shape.n_layers,
                        decoder_attention_heads=shape.n_heads, decoder_ffn_dim=shape.d_ffn,
                        max_target_positions=shape.n_text_ctx, pad_token_id=0, bos_token_id=1, eos_token_id=2,
                        decoder_start_token_id=1)
    hf = WhisperDecoder(cfg).eval()
    sd = hf.state_dict()
    d = shape.d_model

    def T(k):
        return torch.from_numpy(w[k])

    sd["embed_tokens.weight"] = T("decoder/embeddings/weight")
    sd["embed_positions.weight"] = T("decoder/position_encodings/encodings")
    for i in range(shape.n_layers):
        p, q = f"decoder/layer_{i}", f"layers.{i}"
        W, b = T(f"{p}/self_attention/linear_0/weight"), T(f"{p}/self_attention/linear_0/bias")
        sd[f"{q}.self_attn.q_proj.weight"], sd[f"{q}.self_attn.q_proj.bias"] = W[:d], b[:d]
        sd[f"{q}.self_attn.k_proj.weight"] = W[d:2 * d]
        sd[f"{q}.self_attn.v_proj.weight"], sd[f"{q}.self_attn.v_proj.bias"] = W[2 * d:], b[2 * d:]
        sd[f"{q}.self_attn.out_proj.weight"] = T(f"{p}/self_attention/linear_1/weight")
        sd[f"{q}.self_attn.out_proj.bias"] = T(f"{p}/self_attention/linear_1/bias")
        sd[f"{q}.self_attn_layer_norm.weight"] = T(f"{p}/self_attention/layer_norm/gamma")
        sd[f"{q}.self_attn_layer_norm.bias"] = T(f"{p}/self_attention/layer_norm/beta")
        sd[f"{q}.encoder_attn.q_proj.weight"] = T(f"{p}/attention/linear_0/weight")
        sd[f"{q}.encoder_attn.q_proj.bias"] = T(f"{p}/attention/linear_0/bias")
        W, b = T(f"{p}/attention/linear_1/weight"), T(f"{p}/attention/linear_1/bias")
        sd[f"{q}.encoder_attn.k_proj.weight"] = W[:d]
        sd[f"{q}.encoder_attn.v_proj.weight"], sd[f"{q}.encoder_attn.v_proj.bias"] = W[d:], b[d:]
        sd[f"{q}.encoder_attn.out_proj.weight"] = T(f"{p}/attention/linear_2/weight")
        sd[f"{q}.encoder_attn.out_proj.bias"] = T(f"{p}/attention/linear_2/bias")
        sd[f"{q}.encoder_attn_layer_norm.weight"] = T(f"{p}/attention/layer_norm/gamma")
        sd[f"{q}.encoder_attn_layer_norm.bias"] = T(f"{p}/attention/layer_norm/beta")
        sd[f"{q}.final_layer_norm.weight"] = T(f"{p}/ffn/layer_norm/gamma")
        sd[f"{q}.final_layer_norm.bias"] = T(f"{p}/ffn/layer_norm/beta")
        sd[f"{q}.fc1.weight"], sd[f"{q}.fc1.bias"] = T(f"{p}/ffn/linear_0/weight"), T(f"{p}/ffn/linear_0/bias")
        sd[f"{q}.fc2.weight"], sd[f"{q}.fc2.bias"] = T(f"{p}/ffn/linear_1/weight"), T(f"{p}/ffn/linear_1/bias")
    sd["layer_norm.weight"], sd["layer_norm.bias"] = T("decoder/layer_norm/gamma"), T("decoder/layer_norm/beta")
    hf.load_state_dict(sd)
    return hf


def decoder_golden() -> None:
    """tests/golden/decoder_golden.npz (row f1): (a) logits of the oracle decoder and of HF's WhisperDecoder with the
    same weights; (b) the logits-rule masks of the oracle and of HF's WhisperTimeStampLogitsProcessor on random logits
    and token histories; (c) the oracle's greedy ids / decision margins for the seeds the GPU tests use."""
    from transformers.generation.logits_process import WhisperTimeStampLogitsProcessor
    from . import whisper_decoder as wd
    out: dict[str, np.ndarray] = {}
    shape = synth.DEC_SHAPES["micro"]
    tok = synth.WhisperTokens.for_vocab(shape.vocab)
    for tied in (False, True):
        w = synth.decoder_weights(shape, 4321, tied=tied)
        g = torch.Generator().manual_seed(5)
        enc = torch.randn(2, shape.n_audio_ctx, shape.d_model, generator=g)
        toks = torch.randint(0, shape.vocab, (2, 12), generator=g)
        lg = wd.Decoder(w, shape).logits(toks, enc)
        with torch.no_grad():
            h = hf_decoder(shape, w)(input_ids=toks, encoder_hidden_states=enc).last_hidden_state
            lg_hf = h @ torch.from_numpy(w.get("decoder/projection/weight", w["decoder/embeddings/weight"])).T
        err = float((lg - lg_hf).abs().max())
        print("decoder logits oracle vs HF (tied=%s): max abs %.2e of scale %.2f" % (tied, err, float(lg.abs().max())))
        assert err < 5e-5
        key = "tied" if tied else "untied"
        out[f"logits_{key}_tokens"] = toks.numpy()
        out[f"logits_{key}_oracle"] = lg[:, :, ::7].numpy()
        out[f"logits_{key}_hf"] = lg_hf[:, :, ::7].numpy()

    class GC:
        pass
    gc = GC()
    gc.eos_token_id, gc.no_timestamps_token_id, gc.max_initial_timestamp_index = tok.eot, tok.no_timestamps, 50
    P = 3
    proc = WhisperTimeStampLogitsProcessor(gc, begin_index=P)
    rng = np.random.default_rng(0)
    opts = wd.GenerateOptions(suppress_blank=False)
    hist, logit_rows, masks = [], [], []
    for _ in range(80):
        seq = [tok.sot, tok.first_lang, tok.transcribe]
        for _ in range(int(rng.integers(0, 8))):
            seq.append(int(rng.integers(tok.timestamp_begin, shape.vocab)) if rng.random() < 0.4
                       else int(rng.integers(0, tok.eot)))
        logits = (rng.standard_normal(shape.vocab) * 3).astype(np.float32)
        if rng.random() < 0.5:
            logits[tok.timestamp_begin:] += 2.0
        logits = logits.astype(np.float16).astype(np.float32)          # stored as f16 to keep the fixture small
        mine, _ = wd.apply_rules(logits, seq, P, tok, opts, True)
        theirs = proc(torch.tensor([seq]), torch.tensor(logits[None]).clone())[0].numpy()
        assert (np.isneginf(mine) == np.isneginf(theirs)).all(), seq
        hist.append(seq + [-1] * (12 - len(seq)))
        logit_rows.append(logits)
        masks.append(np.packbits(np.isneginf(theirs)))
    out["rules_histories"], out["rules_hf_masks"] = np.array(hist), np.array(masks)
    out["rules_logits"] = np.array(logit_rows).astype(np.float16)

    for name, batch, timestamps, seed, tied in (("micro", 2, True, 32, False), ("micro", 2, False, 23, False),
                                                ("mini", 2, True, 35, False), ("micro", 3, False, 4321, True)):
        shp = synth.DEC_SHAPES[name]
        tk = synth.WhisperTokens.for_vocab(shp.vocab)
        dec = wd.Decoder(synth.decoder_weights(shp, seed, tied=tied), shp, round_weights_bf16=True)
        g = torch.Generator().manual_seed(seed + batch)
        enc = torch.randn(batch, shp.n_audio_ctx, shp.d_model, generator=g).bfloat16()
        prompt = [tk.sot, tk.first_lang + 1, tk.transcribe] + ([] if timestamps else [tk.no_timestamps])
        ref = wd.generate(dec, enc.float(), [prompt] * batch, tk, wd.GenerateOptions(max_length=len(prompt) + 28))
        key = f"greedy_{name}_{seed}"
        out[key + "_ids"] = np.array([r["sequences_ids"] + [-1] * (28 - len(r["sequences_ids"])) for r in ref])
        out[key + "_margins"] = np.array([r["margins"] + [0.0] * (28 - len(r["margins"])) for r in ref], dtype=np.float32)
        out[key + "_score"] = np.array([r["score"] for r in ref], dtype=np.float32)
        out[key + "_nsp"] = np.array([r["no_speech_prob"] for r in ref], dtype=np.float32)
        print(key, [r["sequences_ids"][:6] for r in ref])
    np.savez_compressed(os.path.join(GOLDEN, "decoder_golden.npz"), **out)


def main() -> int:
    if len(sys.argv) > 1 and sys.argv[1] == "decoder":
        decoder_golden()
        return 0
    os.makedirs(GOLDEN, exist_ok=True)
    from transformers import WhisperFeatureExtractor

    mel: dict[str, np.ndarray] = {"frame_pick": FRAME_PICK}
    gens = {"tone": (synth.tone_noise, 0), "chirp": (synth.am_chirp, 1), "gapped": (synth.gapped, 2)}
    for n_mels in (80, 128):
        fe = WhisperFeatureExtractor(feature_size=n_mels)
        for name, (gen, seed) in gens.items():
            x = gen(seed)
            m = logmel.log_mel(x, n_mels)                       # [n_mels, 3001]
            key = f"m{n_mels}_{name}"
            mel[key + "_pick"] = m[:, FRAME_PICK]
            mel[key + "_sha256"] = np.frombuffer(_sha(m).encode(), dtype=np.uint8)
            hf = fe(x, sampling_rate=16000, return_tensors="np")["input_features"][0]
            pick_hf = FRAME_PICK[FRAME_PICK < 2998]
            mel[key + "_hf_pick"] = hf[:, pick_hf].astype(np.float32)
        # short / ragged calls (whole arrays): 1 s, the reference's 0.5-s warm-up length
        # (ref: final_optimized_transcriber.py:188), a length that is not a multiple of the hop, and 185 s / 10
        for n in (16000, 8000, 1234, 29600):
            x = synth.window_signal(7, n)
            mel[f"m{n_mels}_short{n}"] = logmel.log_mel(x, n_mels)
    np.savez_compressed(os.path.join(GOLDEN, "logmel_golden.npz"), **mel)

    enc: dict[str, np.ndarray] = {}
    probe_tokens = {}
    for shape_name, seed in (("micro", 1234), ("tiny", 1234)):
        shape = synth.SHAPES[shape_name]
        w = synth.encoder_weights(shape, seed)
        feats = np.stack([logmel.log_mel_window(synth.window_signal(s), shape.n_mels) for s in (0, 1)])
        out = encoder.encoder_forward(feats, w, shape)
        with torch.no_grad():
            hf_out = hf_encoder(shape, w)(torch.from_numpy(feats)).last_hidden_state
        cmp = encoder.compare(out, hf_out)
        print(shape_name, "oracle vs HF:", cmp)
        assert cmp["max_abs"] < 5e-5, cmp
        enc[f"{shape_name}_out_pick"] = out[:, ::50, ::8].numpy()
        enc[f"{shape_name}_hf_pick"] = hf_out[:, ::50, ::8].numpy()
        probe, toks, margin = GreedyProbe.pick(out, shape.d_model, shape.n_heads)
        print(shape_name, "probe tokens", toks[0, :8].tolist(), "min margin", margin)
        enc[f"{shape_name}_probe_tokens"] = toks.numpy()
        probe_tokens[shape_name] = margin
    np.savez_compressed(os.path.join(GOLDEN, "encoder_golden.npz"), **enc)
    decoder_golden()
    for f in os.listdir(GOLDEN):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
