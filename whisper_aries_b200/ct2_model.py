"""CTranslate2 model directory -> encoder weights.

The reference loads ``Systran/faster-whisper-large-v3`` from ``./models`` (ref: config.py:11,
final_optimized_transcriber.py:172-182): a directory with ``model.bin`` (CTranslate2's variable container),
``config.json`` and ``preprocessor_config.json``.  This module reads that directory on the host and hands the encoder
variables, under the names CTranslate2 stores them by (``encoder/layer_0/self_attention/linear_0/weight`` ...), to
``WhisperModel`` -- so real weights can drive the B200 encoder without going through ctranslate2.

On-disk format of ``model.bin`` (ctranslate2 4.x ``ModelSpec._serialize``, binary version 6), little endian:

    u32 binary_version | str spec_name | u32 spec_revision | u32 n_variables
    n_variables x { str name | u8 rank | u32 dim[rank] | u8 dtype_id | u32 n_bytes | raw bytes }
    u32 n_aliases | n_aliases x { str alias | str variable_name }
    str := u16 length (INCLUDING the terminating NUL) | bytes | NUL
    dtype_id: 0 float32, 1 int8, 2 int16, 3 int32, 4 float16, 5 bfloat16

**[unverified offline]**: neither ctranslate2 nor a converted checkpoint exists in the build container, so the layout
above is restated from the published converter source (``python/ctranslate2/specs/model_spec.py``, ``_serialize``:
``CURRENT_BINARY_VERSION = 6``; C++ ``DataType`` enum order FLOAT32, INT8, INT16, INT32, FLOAT16, BFLOAT16).  It is pinned
by a fixture the test assembles byte by byte with ``struct.pack`` from that description -- independently of
``write_model_bin`` below -- including an int8 + per-row ``weight_scale`` matrix, an int16 + scalar scale matrix, float16
and bfloat16 tensors and an alias (tests/test_ct2_model.py), and by the writer / reader round trip.  int8 weights carry a
per-output-row ``weight_scale`` variable, int16 weights a scalar one (``weight = q / scale``); both are folded here and
everything is returned as float32 (the C ABI rounds to bf16).  The file is memory-mapped and parsed once per
``WhisperModel``: the encoder and decoder loaders share the parse, and a tied output projection (an alias of the
embedding in Whisper checkpoints) is not materialised a second time.
"""
from __future__ import annotations

import json
import mmap
import os
import struct

import numpy as np

from .synthetic import SHAPES, EncoderShape

BINARY_VERSION = 6
_DTYPES = {0: np.dtype("<f4"), 1: np.dtype("i1"), 2: np.dtype("<i2"), 3: np.dtype("<i4"), 4: np.dtype("<f2"),
           5: np.dtype("<u2")}          # 5 = bfloat16, carried as raw u16
_DTYPE_IDS = {"float32": 0, "int8": 1, "int16": 2, "int32": 3, "float16": 4, "bfloat16": 5}


def _bf16_to_f32(raw_u16: np.ndarray) -> np.ndarray:
    return (raw_u16.astype(np.uint32) << 16).view(np.float32)


def _f32_to_bf16(x: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)       # round to nearest even


class _Reader:
    def __init__(self, buf: bytes):
        self.buf, self.pos = buf, 0

    def take(self, fmt: str):
        size = struct.calcsize(fmt)
        if self.pos + size > len(self.buf):
            raise ValueError("model.bin is truncated")
        (v,) = struct.unpack_from("<" + fmt, self.buf, self.pos)
        self.pos += size
        return v

    def string(self) -> str:
        n = self.take("H")
        if n == 0 or self.pos + n > len(self.buf):
            raise ValueError("model.bin: bad string length")
        s = self.buf[self.pos:self.pos + n - 1].decode("utf-8")
        self.pos += n
        return s

    def skip(self, n: int) -> int:
        """Advance past ``n`` payload bytes; returns their offset (the payload is read lazily, straight from the map)."""
        if self.pos + n > len(self.buf):
            raise ValueError("model.bin is truncated")
        at = self.pos
        self.pos += n
        return at


def read_model_bin(path: str, prefix: str = "") -> tuple[dict, dict]:
    """Returns ``({name: ndarray (stored dtype; bfloat16 as raw uint16 tagged in meta)}, meta)`` for every variable
    whose name starts with ``prefix``.  ``meta`` holds spec name / revision, binary version, aliases and the set of
    bfloat16 variable names."""
    with open(path, "rb") as f:
        size = os.fstat(f.fileno()).st_size
        if size < 4:
            raise ValueError("model.bin is truncated")
        buf = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)      # a ~3 GB checkpoint is never copied wholesale
    r = _Reader(buf)
    version = r.take("I")
    if not 1 <= version <= BINARY_VERSION:
        raise ValueError(f"model.bin: unsupported binary version {version}")
    spec = r.string() if version >= 2 else ""
    revision = r.take("I") if version >= 3 else 1
    variables, bf16 = {}, set()
    for _ in range(r.take("I")):
        name = r.string()
        rank = r.take("B")
        dims = [r.take("I") for _ in range(rank)]
        if version >= 4:
            dtype_id = r.take("B")
            n_bytes = r.take("I")
            if dtype_id not in _DTYPES:
                raise ValueError(f"model.bin: variable {name!r} has unknown dtype id {dtype_id}")
            dt = _DTYPES[dtype_id]
        else:                                   # versions < 4 stored the item size instead of a dtype id
            item = r.take("B")
            n_bytes = r.take("I") * item
            dt, dtype_id = {4: np.dtype("<f4"), 2: np.dtype("<i2"), 1: np.dtype("i1")}[item], -1
        n = int(np.prod(dims, dtype=np.int64)) if dims else 1
        if n * dt.itemsize != n_bytes:
            raise ValueError(f"model.bin: variable {name!r}: {n_bytes} bytes for shape {dims} of {dt}")
        at = r.skip(n_bytes)
        if name.startswith(prefix):
            variables[name] = np.frombuffer(buf, dtype=dt, count=n, offset=at).reshape(dims)   # view of the map
            if dtype_id == 5:
                bf16.add(name)
    aliases = {}
    if version >= 3 and r.pos < len(r.buf):
        for _ in range(r.take("I")):
            alias, target = r.string(), r.string()
            aliases[alias] = target
    return variables, {"spec": spec, "revision": revision, "binary_version": version, "aliases": aliases, "bf16": bf16}


def write_model_bin(path: str, variables: dict, spec: str = "WhisperSpec", revision: int = 3, dtypes: dict | None = None,
                    aliases: dict | None = None) -> None:
    """The inverse of ``read_model_bin`` (test infrastructure and export): ``dtypes`` maps a variable name to
    ``"float16"`` / ``"bfloat16"`` / ``"int8"``; int8 weights are quantised per output row and a ``*_scale`` variable
    is added, as the CTranslate2 converter does."""
    dtypes, out = dtypes or {}, []
    for name, arr in variables.items():
        a = np.asarray(arr)
        kind = dtypes.get(name, "float32" if a.dtype.kind == "f" else str(a.dtype))
        if kind == "float32":
            out.append((name, np.ascontiguousarray(a, "<f4"), 0))
        elif kind == "float16":
            out.append((name, np.ascontiguousarray(a, "<f2"), 4))
        elif kind == "bfloat16":
            out.append((name, _f32_to_bf16(a).reshape(a.shape), 5))
        elif kind == "int8":
            a = np.ascontiguousarray(a, np.float32)
            amax = np.abs(a.reshape(a.shape[0], -1)).max(axis=1)
            scale = (127.0 / np.where(amax == 0, 1.0, amax)).astype(np.float32)
            q = np.clip(np.rint(a * scale.reshape(-1, *([1] * (a.ndim - 1)))), -127, 127).astype(np.int8)
            out.append((name, q, 1))
            out.append((name + "_scale", scale, 0))
        else:
            out.append((name, np.ascontiguousarray(a), _DTYPE_IDS[kind]))

    def s(x: str) -> bytes:
        b = x.encode("utf-8")
        return struct.pack("<H", len(b) + 1) + b + b"\0"

    with open(path, "wb") as f:
        f.write(struct.pack("<I", BINARY_VERSION) + s(spec) + struct.pack("<II", revision, len(out)))
        for name, a, dtype_id in out:
            f.write(s(name) + struct.pack("<B", a.ndim) + b"".join(struct.pack("<I", d) for d in a.shape))
            f.write(struct.pack("<BI", dtype_id, a.nbytes) + a.tobytes())
        aliases = aliases or {}
        f.write(struct.pack("<I", len(aliases)) + b"".join(s(k) + s(v) for k, v in aliases.items()))


def _dequantise(name: str, variables: dict, meta: dict) -> np.ndarray:
    a = variables[name]
    if name in meta["bf16"]:
        return _bf16_to_f32(np.ascontiguousarray(a))
    if a.dtype.kind == "i" and name + "_scale" in variables:
        scale = np.asarray(variables[name + "_scale"], np.float32)
        if scale.ndim == 0 or scale.size == 1:
            return a.astype(np.float32) / float(scale.reshape(-1)[0])
        return a.astype(np.float32) / scale.reshape(-1, *([1] * (a.ndim - 1)))
    return a.astype(np.float32)


def encoder_shape_of(weights: dict, name: str = "ct2") -> EncoderShape:
    """Infers (n_mels, d_model, layers, ffn) from the variables; heads = d_model / 64 for every Whisper size."""
    conv1 = weights["encoder/conv1/weight"]
    d, n_mels = int(conv1.shape[0]), int(conv1.shape[1])
    layers = 0
    while f"encoder/layer_{layers}/ffn/linear_0/weight" in weights:
        layers += 1
    if layers == 0:
        raise ValueError("no encoder layers found (encoder/layer_0/...)")
    ffn = int(weights["encoder/layer_0/ffn/linear_0/weight"].shape[0])
    for known in SHAPES.values():
        if (known.n_mels, known.d_model, known.n_layers, known.d_ffn) == (n_mels, d, layers, ffn):
            return known
    if d % 64:
        raise ValueError(f"d_model {d} is not a multiple of the head size 64")
    return EncoderShape(name, n_mels, d, d // 64, layers, ffn)


def parse_model_dir(model_dir: str) -> tuple[dict, dict]:
    """Parses ``model.bin`` ONCE (memory-mapped views) for both loaders below: ``(variables, meta)``."""
    path = os.path.join(model_dir, "model.bin")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path}: not a CTranslate2 model directory")
    return read_model_bin(path)


def _select(parsed: tuple[dict, dict], prefix: str, skip_aliases_of=()) -> tuple[dict, dict]:
    """float32 weights of the variables under ``prefix``; an alias shares its target's array (no second dequantised copy)
    and aliases whose target is in ``skip_aliases_of`` are left out altogether (tied output projection)."""
    variables, meta = parsed
    names = [n for n in variables if n.startswith(prefix) and not n.endswith("_scale") and variables[n].ndim >= 1]
    weights = {n: _dequantise(n, variables, meta) for n in names}
    for alias, target in meta["aliases"].items():
        if alias.startswith(prefix) and target in weights and target not in skip_aliases_of:
            weights[alias] = weights[target]
    return weights, meta


def load_encoder_weights(model_dir: str, parsed=None) -> tuple[EncoderShape, dict, dict]:
    """``(shape, {CT2 variable name: float32 ndarray}, info)`` for the ``encoder/`` variables of a converted model
    directory.  ``info`` carries ``config.json`` / ``preprocessor_config.json`` (when present) and the container's meta
    data; ``preprocessor_config.json``'s ``feature_size`` must agree with conv1's input channels."""
    weights, meta = _select(parsed or parse_model_dir(model_dir), "encoder/")
    shape = encoder_shape_of(weights, os.path.basename(os.path.normpath(model_dir)) or "ct2")
    info = {"meta": {k: v for k, v in meta.items() if k != "bf16"}}
    for fname in ("config.json", "preprocessor_config.json"):
        p = os.path.join(model_dir, fname)
        if os.path.exists(p):
            with open(p) as f:
                info[fname] = json.load(f)
    feat = info.get("preprocessor_config.json", {}).get("feature_size")
    if feat is not None and int(feat) != shape.n_mels:
        raise ValueError(f"preprocessor_config.json feature_size {feat} != conv1 input channels {shape.n_mels}")
    return shape, weights, info


def decoder_shape_of(weights: dict, name: str = "ct2"):
    """Infers the text-decoder dims (vocab, d_model, layers, ffn, positions) from the ``decoder/`` variables (row f1)."""
    from .synthetic import DEC_SHAPES, DecoderShape
    emb = weights["decoder/embeddings/weight"]
    vocab, d = int(emb.shape[0]), int(emb.shape[1])
    layers = 0
    while f"decoder/layer_{layers}/ffn/linear_0/weight" in weights:
        layers += 1
    if layers == 0:
        raise ValueError("no decoder layers found (decoder/layer_0/...)")
    ffn = int(weights["decoder/layer_0/ffn/linear_0/weight"].shape[0])
    n_ctx = int(weights["decoder/position_encodings/encodings"].shape[0])
    for known in DEC_SHAPES.values():
        if (known.vocab, known.d_model, known.n_layers, known.d_ffn, known.n_text_ctx) == (vocab, d, layers, ffn, n_ctx):
            return known
    if d % 64:
        raise ValueError(f"d_model {d} is not a multiple of the head size 64")
    return DecoderShape(name, vocab, d, d // 64, layers, ffn, n_ctx)


def load_decoder_weights(model_dir: str, parsed=None):
    """``(shape, {CT2 variable name: float32 ndarray}, info)`` for the ``decoder/`` variables of a converted model
    directory (embeddings, learned positions, per-layer self-attention / cross-attention / ffn, final LayerNorm and
    the output projection -- an alias of the embedding in Whisper checkpoints).  ``info["suppress_ids"]`` /
    ``info["suppress_ids_begin"]`` come from ``config.json`` (what upstream's ``suppress_tokens=[-1]`` expands to).
    A projection that is an alias of the embedding is NOT returned as a second matrix: the decoder then takes its tied
    path and keeps one copy of the 51866 x 1280 matrix in HBM (``info["tied_projection"]`` says so)."""
    parsed = parsed or parse_model_dir(model_dir)
    tied = parsed[1]["aliases"].get("decoder/projection/weight") == "decoder/embeddings/weight"
    weights, meta = _select(parsed, "decoder/", skip_aliases_of=("decoder/embeddings/weight",) if tied else ())
    shape = decoder_shape_of(weights, os.path.basename(os.path.normpath(model_dir)) or "ct2")
    info = {"meta": {k: v for k, v in meta.items() if k != "bf16"}, "suppress_ids": [], "suppress_ids_begin": [],
            "tied_projection": tied}
    p = os.path.join(model_dir, "config.json")
    if os.path.exists(p):
        with open(p) as f:
            cfg = json.load(f)
        info["config.json"] = cfg
        info["suppress_ids"] = [int(t) for t in cfg.get("suppress_ids", [])]
        info["suppress_ids_begin"] = [int(t) for t in cfg.get("suppress_ids_begin", [])]
    return shape, weights, info
