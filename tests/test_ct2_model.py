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
    assert np.array_equal(got["decoder/projection/weight"], got["decoder/embeddings/weight"])     # the alias
    with pytest.raises((KeyError, ValueError)):
        enc_only = tmp_path / "enc"
        enc_only.mkdir()
        ct2_model.write_model_bin(str(enc_only / "model.bin"), dict(osynth.encoder_weights(osynth.SHAPES["micro"], 7)))
        ct2_model.load_decoder_weights(str(enc_only))
