"""CTranslate2 model-directory loader: container round trips on the CPU, and on the GPU the encoder
built from a directory equals the one built from the same weights passed as a dict."""
import dataclasses
import json
import os
import struct

import numpy as np
import pytest

from oracle import synth as osynth
from whisper_aries_b200 import ct2_model


def micro_dir(tmp_path, dtypes=None, extra=None):
    shape = osynth.SHAPES["micro"]
    w = osynth.encoder_weights(shape, 7)
    variables = dict(w)
    variables["decoder/embeddings/weight"] = np.ones((5, 128), np.float32)        # must be skipped by the prefix
    variables.update(extra or {})
    d = tmp_path / "model"
    d.mkdir()
    ct2_model.write_model_bin(str(d / "model.bin"), variables, dtypes=dtypes or {},
                              aliases={"encoder/alias_of_conv1_bias": "encoder/conv1/bias"})
    (d / "config.json").write_text(json.dumps({"lang_ids": [1, 2, 3]}))
    (d / "preprocessor_config.json").write_text(json.dumps({"feature_size": shape.n_mels, "n_fft": 400}))
    return str(d), shape, w


def test_float32_round_trip_is_exact(tmp_path):
    path, shape, w = micro_dir(tmp_path)
    got_shape, got, info = ct2_model.load_encoder_weights(path)
    assert dataclasses.astuple(got_shape) == dataclasses.astuple(shape)
    assert info["meta"]["spec"] == "WhisperSpec" and info["meta"]["binary_version"] == 6
    assert info["preprocessor_config.json"]["feature_size"] == shape.n_mels and "config.json" in info
    assert set(w) <= set(got) and not any(k.startswith("decoder/") for k in got)
    for k, v in w.items():
        assert got[k].dtype == np.float32 and np.array_equal(got[k], v), k
    assert np.array_equal(got["encoder/alias_of_conv1_bias"], w["encoder/conv1/bias"])


def test_half_bfloat16_and_int8_containers(tmp_path):
    shape = osynth.SHAPES["micro"]
    names = [k for k in osynth.encoder_weights(shape, 7) if k.endswith("/weight")]
    for kind, tol in (("float16", 2 ** -11), ("bfloat16", 2 ** -8), ("int8", 1 / 127)):
        sub = tmp_path / kind
        sub.mkdir()
        path, _, w = micro_dir(sub, dtypes={n: kind for n in names})
        _, got, _ = ct2_model.load_encoder_weights(path)
        for n in names:
            ref = w[n]
            scale = np.abs(ref).max() if kind != "int8" else np.abs(ref.reshape(ref.shape[0], -1)).max(axis=1).reshape(
                -1, *([1] * (ref.ndim - 1)))
            assert np.all(np.abs(got[n] - ref) <= tol * scale * 1.01 + 1e-12), (kind, n)
        assert not any(k.endswith("_scale") for k in got)
        assert np.array_equal(got["encoder/conv1/bias"], w["encoder/conv1/bias"])       # untouched f32 variables


def test_shape_inference_and_errors(tmp_path):
    path, shape, w = micro_dir(tmp_path)
    assert dataclasses.astuple(ct2_model.encoder_shape_of(w)) == dataclasses.astuple(shape)
    odd = dict(w)
    odd["encoder/conv1/weight"] = np.zeros((192, 80, 3), np.float32)
    odd_shape = ct2_model.encoder_shape_of(odd, "custom")
    assert (odd_shape.d_model, odd_shape.n_heads, odd_shape.name) == (192, 3, "custom")
    with pytest.raises(FileNotFoundError):
        ct2_model.load_encoder_weights(str(tmp_path / "nope"))
    raw = open(os.path.join(path, "model.bin"), "rb").read()
    trunc = tmp_path / "trunc"
    trunc.mkdir()
    (trunc / "model.bin").write_bytes(raw[: len(raw) // 2])
    with pytest.raises(ValueError, match="truncated"):
        ct2_model.load_encoder_weights(str(trunc))
    bad = tmp_path / "badver"
    bad.mkdir()
    (bad / "model.bin").write_bytes(struct.pack("<I", 99) + raw[4:])
    with pytest.raises(ValueError, match="binary version"):
        ct2_model.load_encoder_weights(str(bad))
    wrong = tmp_path / "wrongfeat"
    wrong.mkdir()
    (wrong / "model.bin").write_bytes(raw)
    (wrong / "preprocessor_config.json").write_text(json.dumps({"feature_size": 128}))
    with pytest.raises(ValueError, match="feature_size"):
        ct2_model.load_encoder_weights(str(wrong))


@pytest.mark.gpu
def test_model_from_directory_matches_model_from_dict(tmp_path):
    import torch
    from whisper_aries_b200 import WhisperModel
    path, shape, w = micro_dir(tmp_path)
    from_dir = WhisperModel(path, device="cuda", device_index=0)
    from_dict = WhisperModel("micro", w, device="cuda", device_index=0)
    assert dataclasses.astuple(from_dir.shape) == dataclasses.astuple(shape) and from_dir.model_info["meta"]["spec"] == "WhisperSpec"
    pcm = torch.from_numpy(osynth.batch_signals(2, 3)).cuda()
    assert torch.equal(from_dir.encode_audio(pcm), from_dict.encode_audio(pcm))
    with pytest.raises(ValueError, match="CTranslate2 model directory"):
        WhisperModel("large-v3", device="cuda")


def test_decoder_variables_round_trip(tmp_path):
    """Row f1: the decoder/ variables of the same container, with the output projection stored as an alias of the
    embedding (as Whisper checkpoints do) and the suppress list taken from config.json."""
    shape = osynth.DEC_SHAPES["micro"]
    w = osynth.decoder_weights(shape, 3, tied=True)
    variables = dict(osynth.encoder_weights(osynth.SHAPES["micro"], 7))
    variables.update(w)
    d = tmp_path / "model"
    d.mkdir()
    ct2_model.write_model_bin(str(d / "model.bin"), variables, dtypes={"decoder/embeddings/weight": "float16"},
                              aliases={"decoder/projection/weight": "decoder/embeddings/weight"})
    (d / "config.json").write_text(json.dumps({"suppress_ids": [1, 2, 7], "suppress_ids_begin": [220, 50257]}))
    got_shape, got, info = ct2_model.load_decoder_weights(str(d))
    assert dataclasses.astuple(got_shape) == dataclasses.astuple(shape)
    assert info["suppress_ids"] == [1, 2, 7] and info["suppress_ids_begin"] == [220, 50257]
    assert not any(k.startswith("encoder/") for k in got)
    for k, v in w.items():
        if k == "decoder/embeddings/weight":
            assert np.abs(got[k] - v).max() <= 2 ** -11 * np.abs(v).max()
        else:
            assert np.array_equal(got[k], v), k
    # the tied projection is not materialised a second time: the decoder takes its tied path (one matrix in HBM)
    assert "decoder/projection/weight" not in got and info["tied_projection"] is True
    with pytest.raises((KeyError, ValueError)):
        enc_only = tmp_path / "enc"
        enc_only.mkdir()
        ct2_model.write_model_bin(str(enc_only / "model.bin"), dict(osynth.encoder_weights(osynth.SHAPES["micro"], 7)))
        ct2_model.load_decoder_weights(str(enc_only))


def _ct2_fixture_bytes(entries, aliases, spec=b"WhisperSpec", revision=3, version=6):
    """A ``model.bin`` assembled byte by byte from CTranslate2's documented serialisation -- NOT through
    ``ct2_model.write_model_bin`` -- following ``ModelSpec._serialize`` (python/ctranslate2/specs/model_spec.py, 4.x):

        model.write(struct.pack("I", CURRENT_BINARY_VERSION))          # 6
        _write_string(self.name); model.write(struct.pack("I", self.revision))
        model.write(struct.pack("I", len(variables)))
        for name, value in variables:
            _write_string(name); model.write(struct.pack("B", len(value.shape)))
            for dim in value.shape: model.write(struct.pack("I", dim))
            model.write(struct.pack("B", value.dtype id)); model.write(struct.pack("I", value.num_bytes()))
            model.write(value.to_bytes())
        model.write(struct.pack("I", len(aliases)))
        for alias, variable_name in aliases: _write_string(alias); _write_string(variable_name)
        _write_string(s) = pack("H", len(s) + 1) + s.encode("utf-8") + pack("B", 0)

    with the C++ ``DataType`` ids FLOAT32 0, INT8 1, INT16 2, INT32 3, FLOAT16 4, BFLOAT16 5 (include/ctranslate2/types.h).
    ``entries``: (name bytes, shape tuple, dtype id, payload bytes)."""
    def wstr(b):
        return struct.pack("<H", len(b) + 1) + b + b"\x00"
    out = bytearray()
    out += struct.pack("<I", version) + wstr(spec) + struct.pack("<I", revision) + struct.pack("<I", len(entries))
    for name, shape, dtype_id, payload in entries:
        out += wstr(name) + struct.pack("<B", len(shape))
        for dim in shape:
            out += struct.pack("<I", dim)
        out += struct.pack("<B", dtype_id) + struct.pack("<I", len(payload)) + payload
    out += struct.pack("<I", len(aliases))
    for alias, target in aliases:
        out += wstr(alias) + wstr(target)
    return bytes(out)


def test_reader_against_a_fixture_built_from_the_published_serialisation(tmp_path):
    """Row f2 pinned against something other than its own writer: every byte below comes from struct.pack calls laid
    out as CTranslate2's converter writes them; expected values are computed by hand here (int8 / scale, int16 / scalar
    scale, IEEE half and bfloat16 bit patterns)."""
    # float32 vector
    f32 = struct.pack("<4f", 1.5, -2.25, 0.0, 3.0e-3)
    # int8 matrix [2, 3] with per-row scales (weight = q / scale): rows scaled by 127 / amax
    q8 = struct.pack("<6b", 127, -64, 0, -127, 1, 32)
    s8 = struct.pack("<2f", 254.0, 63.5)
    # int16 matrix [1, 2] with ONE scalar scale (rank-0 variable)
    q16 = struct.pack("<2h", 1000, -32767)
    s16 = struct.pack("<f", 2000.0)
    # float16: 0x3C00 = 1.0, 0xC000 = -2.0, 0x3555 ~ 0.333251953125; bfloat16: 0x3F80 = 1.0, 0xBF00 = -0.5, 0x4049 = 3.140625
    f16 = struct.pack("<3H", 0x3C00, 0xC000, 0x3555)
    bf16 = struct.pack("<3H", 0x3F80, 0xBF00, 0x4049)
    entries = [
        (b"encoder/conv1/bias", (4,), 0, f32),
        (b"encoder/layer_0/ffn/linear_0/weight", (2, 3), 1, q8),
        (b"encoder/layer_0/ffn/linear_0/weight_scale", (2,), 0, s8),
        (b"encoder/layer_0/ffn/linear_1/weight", (1, 2), 2, q16),
        (b"encoder/layer_0/ffn/linear_1/weight_scale", (), 0, s16),
        (b"encoder/layer_norm/gamma", (3,), 4, f16),
        (b"decoder/layer_norm/beta", (3,), 5, bf16),
        (b"decoder/embeddings/weight", (1, 3), 4, f16),
    ]
    blob = _ct2_fixture_bytes(entries, [(b"decoder/projection/weight", b"decoder/embeddings/weight")])
    # spot-check the framing itself: version, "WhisperSpec\0" with its length prefix, revision, count
    assert blob[:4] == b"\x06\x00\x00\x00" and blob[4:6] == b"\x0c\x00" and blob[6:18] == b"WhisperSpec\x00"
    assert blob[18:22] == b"\x03\x00\x00\x00" and blob[22:26] == b"\x08\x00\x00\x00"
    path = tmp_path / "model.bin"
    path.write_bytes(blob)
    variables, meta = ct2_model.read_model_bin(str(path))
    assert meta["spec"] == "WhisperSpec" and meta["revision"] == 3 and meta["binary_version"] == 6
    assert meta["aliases"] == {"decoder/projection/weight": "decoder/embeddings/weight"}
    assert variables["encoder/layer_0/ffn/linear_0/weight"].dtype == np.int8
    assert variables["encoder/layer_0/ffn/linear_1/weight_scale"].shape == ()
    enc, _ = ct2_model._select((variables, meta), "encoder/")
    assert np.array_equal(enc["encoder/conv1/bias"], np.array([1.5, -2.25, 0.0, 3.0e-3], np.float32))
    assert np.allclose(enc["encoder/layer_0/ffn/linear_0/weight"],
                       np.array([[127 / 254.0, -64 / 254.0, 0.0], [-127 / 63.5, 1 / 63.5, 32 / 63.5]]), rtol=0, atol=1e-7)
    assert np.allclose(enc["encoder/layer_0/ffn/linear_1/weight"], np.array([[0.5, -32767 / 2000.0]]), rtol=0, atol=1e-6)
    assert np.array_equal(enc["encoder/layer_norm/gamma"], np.array([1.0, -2.0, 0.333251953125], np.float32))
    assert not any(k.endswith("_scale") for k in enc)
    dec, _ = ct2_model._select((variables, meta), "decoder/")
    assert np.array_equal(dec["decoder/layer_norm/beta"], np.array([1.0, -0.5, 3.140625], np.float32))
    assert dec["decoder/projection/weight"] is dec["decoder/embeddings/weight"]          # alias shares the array
    # and the package's own writer emits the same bytes for the same float32 / float16 content (writer pinned too)
    own = tmp_path / "own.bin"
    ct2_model.write_model_bin(str(own), {"encoder/conv1/bias": np.array([1.5, -2.25, 0.0, 3.0e-3], np.float32),
                                         "encoder/layer_norm/gamma": np.array([1.0, -2.0, 0.333251953125], np.float32)},
                              dtypes={"encoder/layer_norm/gamma": "float16"}, aliases={"a/b": "encoder/conv1/bias"})
    want = _ct2_fixture_bytes([(b"encoder/conv1/bias", (4,), 0, f32), (b"encoder/layer_norm/gamma", (3,), 4, f16)],
                              [(b"a/b", b"encoder/conv1/bias")])
    assert own.read_bytes() == want
    # truncation anywhere inside the variable table is reported, not mis-parsed
    (tmp_path / "cut.bin").write_bytes(blob[:60])
    with pytest.raises(ValueError):
        ct2_model.read_model_bin(str(tmp_path / "cut.bin"))
