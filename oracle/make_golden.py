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


def main() -> int:
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
    for f in os.listdir(GOLDEN):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
