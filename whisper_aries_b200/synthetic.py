"""Deterministic synthetic audio and random-init encoder weights.

No dataset or checkpoint exists offline, so bench.py / smoke() / the tests feed the kernels with these generators.
Weight names follow CTranslate2's Whisper variables ("encoder/layer_3/ffn/linear_0/weight", ...), i.e. what a
model.bin loader would hand to ``WhisperEncoder``."""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

SAMPLE_RATE = 16000
N_SAMPLES = 480000


@dataclass(frozen=True)
class EncoderShape:
    name: str
    n_mels: int
    d_model: int
    n_heads: int
    n_layers: int
    d_ffn: int
    n_ctx: int = 1500

    @property
    def flops_per_window(self) -> float:
        """2 * MACs of the conv stem + all layers for one 30-s window."""
        d, t = self.d_model, self.n_ctx
        conv = 2 * self.n_mels * 3 * d * 3000 + 2 * d * 3 * d * 1500
        layer = 8 * t * d * d + 4 * t * t * d + 4 * t * d * self.d_ffn
        return float(conv + self.n_layers * layer)


SHAPES = {
    "tiny": EncoderShape("tiny", 80, 384, 6, 4, 1536),
    "medium": EncoderShape("medium", 80, 1024, 16, 24, 4096),
    "large-v3": EncoderShape("large-v3", 128, 1280, 20, 32, 5120),
    # not a Whisper size: the smallest shape every kernel tiling accepts, for fast parity tests
    "micro": EncoderShape("micro", 80, 128, 2, 2, 512),
}


def tone_noise(seed: int, n: int = N_SAMPLES) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(n) / SAMPLE_RATE
    return (0.3 * np.sin(2 * np.pi * 220.0 * t) + 0.05 * rng.standard_normal(n)).astype(np.float32)


def am_chirp(seed: int, n: int = N_SAMPLES) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(n) / SAMPLE_RATE
    car = np.sin(2 * np.pi * (200.0 * t + 40.0 * t * t))
    env = 0.5 + 0.5 * np.sin(2 * np.pi * 3.0 * t)
    return (0.3 * car * env + 1e-3 * rng.standard_normal(n)).astype(np.float32)


def gapped(seed: int, n: int = N_SAMPLES) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = tone_noise(seed + 1000, n)
    seg = SAMPLE_RATE // 2
    n_seg = (n + seg - 1) // seg
    silent = rng.random(n_seg) < 0.3
    mask = np.repeat(~silent, seg)[:n]
    return (x * mask).astype(np.float32)


def window_signal(seed: int, n: int = N_SAMPLES) -> np.ndarray:
    kind = seed % 3
    if kind == 0:
        return tone_noise(seed, n)
    if kind == 1:
        return am_chirp(seed, n)
    return gapped(seed, n)


def batch_signals(n_windows: int, first_seed: int = 0, n: int = N_SAMPLES) -> np.ndarray:
    return np.stack([window_signal(first_seed + i, n) for i in range(n_windows)])


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> np.ndarray:
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = np.exp(-inc * np.arange(channels // 2))
    ang = np.arange(length)[:, None] * inv[None, :]
    return np.concatenate([np.sin(ang), np.cos(ang)], axis=1).astype(np.float32)


def encoder_weights(shape: EncoderShape, seed: int = 1234) -> dict[str, np.ndarray]:
    """N(0, 0.02^2) weights and biases (key bias zero), LayerNorm gamma = 1 + N(0, 0.02^2), sinusoidal positions."""
    rng = np.random.default_rng(seed)
    d, f = shape.d_model, shape.d_ffn

    def nrm(*s):
        return (0.02 * rng.standard_normal(s)).astype(np.float32)

    w: dict[str, np.ndarray] = {}
    w["encoder/conv1/weight"] = nrm(d, shape.n_mels, 3)
    w["encoder/conv1/bias"] = nrm(d)
    w["encoder/conv2/weight"] = nrm(d, d, 3)
    w["encoder/conv2/bias"] = nrm(d)
    w["encoder/position_encodings/encodings"] = sinusoids(shape.n_ctx, d)
    for i in range(shape.n_layers):
        p = f"encoder/layer_{i}"
        w[f"{p}/self_attention/layer_norm/gamma"] = (1.0 + nrm(d)).astype(np.float32)
        w[f"{p}/self_attention/layer_norm/beta"] = nrm(d)
        w[f"{p}/self_attention/linear_0/weight"] = nrm(3 * d, d)
        b = nrm(3 * d)
        b[d:2 * d] = 0.0
        w[f"{p}/self_attention/linear_0/bias"] = b
        w[f"{p}/self_attention/linear_1/weight"] = nrm(d, d)
        w[f"{p}/self_attention/linear_1/bias"] = nrm(d)
        w[f"{p}/ffn/layer_norm/gamma"] = (1.0 + nrm(d)).astype(np.float32)
        w[f"{p}/ffn/layer_norm/beta"] = nrm(d)
        w[f"{p}/ffn/linear_0/weight"] = nrm(f, d)
        w[f"{p}/ffn/linear_0/bias"] = nrm(f)
        w[f"{p}/ffn/linear_1/weight"] = nrm(d, f)
        w[f"{p}/ffn/linear_1/bias"] = nrm(d)
    w["encoder/layer_norm/gamma"] = (1.0 + nrm(d)).astype(np.float32)
    w["encoder/layer_norm/beta"] = nrm(d)
    return w


# ---------------------------------------------------------------------------------------------- decoder (row f1)
@dataclass(frozen=True)
class DecoderShape:
    """Text-decoder dims of a Whisper size (OpenAI dims; vocab 51866 for large-v3, 51865 for the other multilingual sizes)."""
    name: str
    vocab: int
    d_model: int
    n_heads: int
    n_layers: int
    d_ffn: int
    n_text_ctx: int = 448
    n_audio_ctx: int = 1500


@dataclass(frozen=True)
class WhisperTokens:
    """Special-token ids the decoding rules need (tokenizer.json of the converted model; large-v3 values by default)."""
    eot: int = 50257
    sot: int = 50258
    transcribe: int = 50360
    translate: int = 50359
    sot_lm: int = 50361
    sot_prev: int = 50362
    no_speech: int = 50363
    no_timestamps: int = 50364
    timestamp_begin: int = 50365
    blank: int = 220          # " "
    first_lang: int = 50259

    @classmethod
    def for_vocab(cls, vocab: int) -> "WhisperTokens":
        if vocab == 51866:
            return cls()
        if vocab == 51865:      # multilingual sizes before large-v3: one language token fewer
            return cls(50257, 50258, 50359, 50358, 50360, 50361, 50362, 50363, 50364, 220, 50259)
        # synthetic vocabularies (tests): the last 128 ids are the special block, timestamps take what follows
        b = vocab - 128
        return cls(b, b + 1, b + 5, b + 4, b + 6, b + 7, b + 8, b + 9, b + 10, 7, b + 2)


DEC_SHAPES = {
    "tiny": DecoderShape("tiny", 51865, 384, 6, 4, 1536),
    "medium": DecoderShape("medium", 51865, 1024, 16, 24, 4096),
    "large-v3": DecoderShape("large-v3", 51866, 1280, 20, 32, 5120),
    # not Whisper sizes: small shapes every kernel tiling accepts (vocab deliberately not a multiple of 128)
    "micro": DecoderShape("micro", 1000, 128, 2, 2, 512),
    "mini": DecoderShape("mini", 3000, 384, 6, 3, 1536),
}


def decoder_weights(shape: DecoderShape, seed: int = 4321, logit_scale: float = 0.15, tied: bool = False,
                    cross_gain: float = 6.0) -> dict[str, np.ndarray]:
    """Random-init decoder weights under CTranslate2's Whisper variable names [unverified offline] (f32 numpy).

    Matrices ~ N(0, 0.02^2) except the token embedding (= tied output projection) ~ N(0, logit_scale^2) so that
    greedy top-1/top-2 margins are well above bf16 noise; biases ~ N(0, 0.02^2) with the key slices zero (Whisper's
    k_proj has no bias); LayerNorm gamma = 1 + N(0, 0.02^2); learned positions ~ N(0, 0.02^2).

    Real Whisper ties the output projection to the embedding; with RANDOM weights a tied head just repeats the previous
    token and ignores the audio, which would make "identical token ids" a vacuous test.  So by default an untied
    ``decoder/projection/weight`` is added and the cross-attention path is amplified (``cross_gain``) so that the
    decoded ids are diverse and depend on the encoder output; ``tied=True`` gives the Whisper layout."""
    rng = np.random.default_rng(seed)
    d, f = shape.d_model, shape.d_ffn

    def nrm(*s, sc=0.02):
        return (sc * rng.standard_normal(s)).astype(np.float32)

    w: dict[str, np.ndarray] = {}
    w["decoder/embeddings/weight"] = nrm(shape.vocab, d, sc=logit_scale)
    w["decoder/position_encodings/encodings"] = nrm(shape.n_text_ctx, d)
    for i in range(shape.n_layers):
        p = f"decoder/layer_{i}"
        for blk in ("self_attention", "attention", "ffn"):
            w[f"{p}/{blk}/layer_norm/gamma"] = (1.0 + nrm(d)).astype(np.float32)
            w[f"{p}/{blk}/layer_norm/beta"] = nrm(d)
        w[f"{p}/self_attention/linear_0/weight"] = nrm(3 * d, d, sc=0.05)
        b = nrm(3 * d)
        b[d:2 * d] = 0.0
        w[f"{p}/self_attention/linear_0/bias"] = b
        w[f"{p}/self_attention/linear_1/weight"] = nrm(d, d, sc=0.05)
        w[f"{p}/self_attention/linear_1/bias"] = nrm(d)
        w[f"{p}/attention/linear_0/weight"] = nrm(d, d, sc=0.05)            # cross-attention query
        w[f"{p}/attention/linear_0/bias"] = nrm(d)
        w[f"{p}/attention/linear_1/weight"] = nrm(2 * d, d, sc=0.05)        # cross-attention key | value
        b = nrm(2 * d)
        b[:d] = 0.0
        w[f"{p}/attention/linear_1/bias"] = b
        w[f"{p}/attention/linear_2/weight"] = nrm(d, d, sc=0.05)
        w[f"{p}/attention/linear_2/bias"] = nrm(d)
        w[f"{p}/ffn/linear_0/weight"] = nrm(f, d, sc=0.05)
        w[f"{p}/ffn/linear_0/bias"] = nrm(f)
        w[f"{p}/ffn/linear_1/weight"] = nrm(d, f, sc=0.03)
        w[f"{p}/ffn/linear_1/bias"] = nrm(d)
    w["decoder/layer_norm/gamma"] = (1.0 + nrm(d)).astype(np.float32)
    w["decoder/layer_norm/beta"] = nrm(d)
    if not tied:
        rng2 = np.random.default_rng(seed + 99)
        w["decoder/projection/weight"] = (logit_scale * rng2.standard_normal((shape.vocab, d))).astype(np.float32)
        for i in range(shape.n_layers):
            p = f"decoder/layer_{i}/attention"
            w[f"{p}/linear_0/weight"] = w[f"{p}/linear_0/weight"] * 2.0
            w[f"{p}/linear_1/weight"] = w[f"{p}/linear_1/weight"] * 2.0
            w[f"{p}/linear_2/weight"] = w[f"{p}/linear_2/weight"] * np.float32(cross_gain)
    return w


def decoder_weights_fast(shape: DecoderShape, seed: int = 0) -> dict[str, np.ndarray]:
    """Whisper-layout (tied) random decoder weights drawn with torch's multi-threaded generator: 0.8 G parameters for
    large-v3 in a few seconds.  Benchmarks only -- parity tests use ``decoder_weights`` (numpy, shared with the oracle)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    d, f = shape.d_model, shape.d_ffn

    def nrm(*s, sc=0.02):
        return (torch.randn(*s, generator=g) * sc).numpy()

    w = {"decoder/embeddings/weight": nrm(shape.vocab, d, sc=0.05),
         "decoder/position_encodings/encodings": nrm(shape.n_text_ctx, d)}
    for i in range(shape.n_layers):
        p = f"decoder/layer_{i}"
        for blk in ("self_attention", "attention", "ffn"):
            w[f"{p}/{blk}/layer_norm/gamma"] = 1.0 + nrm(d)
            w[f"{p}/{blk}/layer_norm/beta"] = nrm(d)
        for name, n_out, n_in in (("self_attention/linear_0", 3 * d, d), ("self_attention/linear_1", d, d),
                                  ("attention/linear_0", d, d), ("attention/linear_1", 2 * d, d),
                                  ("attention/linear_2", d, d), ("ffn/linear_0", f, d), ("ffn/linear_1", d, f)):
            w[f"{p}/{name}/weight"] = nrm(n_out, n_in)
            w[f"{p}/{name}/bias"] = nrm(n_out)
    w["decoder/layer_norm/gamma"] = 1.0 + nrm(d)
    w["decoder/layer_norm/beta"] = nrm(d)
    return w
