"""fp32 torch-CPU restatement of CTranslate2 4.6.0 ``layers::WhisperEncoder`` (oracle; test infrastructure).

Follows SURVEY.md section 8 rows a-7 / a-8:
  conv1(k3,s1,p1) -> GELU(erf) -> conv2(k3,s2,p1) -> GELU(erf) -> [B,T,d] -> + position_encodings[:T]
  -> N x pre-norm layer { x += O(softmax(q k^T) v) ; x += W2 gelu(W1 LN(x)) } -> LayerNorm (eps 1e-5)
  with q = (LN(x) Wq + bq) * head_dim^-0.5, k without bias, no mask, no dropout.

The reference reaches it through ``model.transcribe`` -> ``WhisperModel.encode`` ->
``ctranslate2.models.Whisper.encode`` (ref: final_optimized_transcriber.py:326; large-v3 forced at
conversation_transcriber.py:72).  Weight names are CTranslate2's Whisper variable names.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .synth import EncoderShape

LN_EPS = 1e-5


def _t(w, key, dtype=torch.float32):
    v = w[key]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v)).to(dtype)


def check_features(features: np.ndarray | torch.Tensor, shape: EncoderShape) -> None:
    """Row a-7 input validation: rank 3, dim1 == n_mels, dim2 <= 3000 else ValueError."""
    if features.ndim != 3 or features.shape[1] != shape.n_mels or features.shape[2] > 3000:
        raise ValueError(
            f"Invalid input features shape: expected an input with shape (batch, {shape.n_mels}, <=3000), "
            f"but got {tuple(features.shape)}")


def build_int8_linears(w: dict, shape: EncoderShape) -> dict:
    """Dynamic-int8 (per-tensor activation scale, int8 weights, fbgemm / x86 engine) stand-ins for the four Linear
    layers of every encoder layer -- the closest torch-CPU equivalent of the reference's ``compute_type="int8"``
    (ref: final_optimized_transcriber.py:205; CT2 quantises the same four GEMMs and keeps everything else f32).
    Used ONLY by the CPU-baseline legs of bench.py; parity checks always use the f32 path."""
    import warnings
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(shape.n_layers):
            for key in (f"encoder/layer_{i}/self_attention/linear_0", f"encoder/layer_{i}/self_attention/linear_1",
                        f"encoder/layer_{i}/ffn/linear_0", f"encoder/layer_{i}/ffn/linear_1"):
                wt, bs = _t(w, key + "/weight"), _t(w, key + "/bias")
                lin = torch.nn.Linear(wt.shape[1], wt.shape[0])
                lin.weight.data, lin.bias.data = wt, bs
                out[key] = torch.ao.quantization.quantize_dynamic(torch.nn.Sequential(lin), {torch.nn.Linear},
                                                                  dtype=torch.qint8)[0]
    return out


@torch.no_grad()
def encoder_forward(features, w: dict, shape: EncoderShape, *, return_layers: bool = False,
                    round_weights_bf16: bool = False, int8_linears: dict | None = None):
    """features f32 [B, n_mels, 3000] -> f32 [B, 1500, d].

    ``round_weights_bf16`` rounds weights (not activations) to bf16 first, to separate the error of
    bf16 weight storage from the error of bf16 arithmetic when judging the CUDA path.
    ``int8_linears`` (from ``build_int8_linears``) swaps the Linear layers for dynamic-int8 ones: CPU baseline only.
    """
    x = features if isinstance(features, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(features))
    x = x.to(torch.float32)
    if x.ndim == 2:
        x = x[None]
    check_features(x, shape)

    def W(key):
        t = _t(w, key)
        if round_weights_bf16 and t.ndim >= 2 and "position" not in key:
            t = t.to(torch.bfloat16).to(torch.float32)
        return t

    def linear(inp, key):
        if int8_linears is not None:
            return int8_linears[key](inp)
        return F.linear(inp, W(key + "/weight"), W(key + "/bias"))

    d, h = shape.d_model, shape.n_heads
    hd = d // h
    x = F.gelu(F.conv1d(x, W("encoder/conv1/weight"), W("encoder/conv1/bias"), stride=1, padding=1))
    x = F.gelu(F.conv1d(x, W("encoder/conv2/weight"), W("encoder/conv2/bias"), stride=2, padding=1))
    x = x.transpose(1, 2)                                             # [B, T, d]
    T = x.shape[1]
    x = x + W("encoder/position_encodings/encodings")[:T]
    layers = []
    for i in range(shape.n_layers):
        p = f"encoder/layer_{i}"
        y = F.layer_norm(x, (d,), W(f"{p}/self_attention/layer_norm/gamma"),
                         W(f"{p}/self_attention/layer_norm/beta"), LN_EPS)
        qkv = linear(y, f"{p}/self_attention/linear_0")
        q, k, v = qkv.split(d, dim=-1)
        B = x.shape[0]
        q = q.view(B, T, h, hd).transpose(1, 2) * (hd ** -0.5)
        k = k.view(B, T, h, hd).transpose(1, 2)
        v = v.view(B, T, h, hd).transpose(1, 2)
        att = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
        ctx = (att @ v).transpose(1, 2).reshape(B, T, d)
        x = x + linear(ctx, f"{p}/self_attention/linear_1")
        y = F.layer_norm(x, (d,), W(f"{p}/ffn/layer_norm/gamma"), W(f"{p}/ffn/layer_norm/beta"), LN_EPS)
        y = F.gelu(linear(y, f"{p}/ffn/linear_0"))
        x = x + linear(y, f"{p}/ffn/linear_1")
        if return_layers:
            layers.append(x.clone())
    out = F.layer_norm(x, (d,), W("encoder/layer_norm/gamma"), W("encoder/layer_norm/beta"), LN_EPS)
    return (out, layers) if return_layers else out


def compare(out: torch.Tensor, ref: torch.Tensor) -> dict:
    """max-abs error and cosine similarity (whole tensor and worst row) — the numbers parity tests state."""
    a = out.detach().to(torch.float32).reshape(-1, out.shape[-1])
    b = ref.detach().to(torch.float32).reshape(-1, ref.shape[-1])
    row_cos = F.cosine_similarity(a, b, dim=-1)
    return {
        "max_abs": float((a - b).abs().max()),
        "ref_max_abs": float(b.abs().max()),
        "cosine": float(F.cosine_similarity(a.reshape(1, -1), b.reshape(1, -1)).item()),
        "min_row_cosine": float(row_cos.min()),
    }
