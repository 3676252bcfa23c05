"""CPU: the fp32 torch encoder oracle against golden vectors (own + HF transformers' WhisperEncoder)."""
import os

import numpy as np
import pytest
import torch

from oracle import encoder, logmel, synth
from oracle.decoder import GreedyProbe


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "encoder_golden.npz"))


@pytest.mark.parametrize("shape_name", ["micro", "tiny"])
def test_encoder_matches_golden_and_hf(gold, shape_name):
    shape = synth.SHAPES[shape_name]
    w = synth.encoder_weights(shape, 1234)
    feats = np.stack([logmel.log_mel_window(synth.window_signal(s), shape.n_mels) for s in (0, 1)])
    out = encoder.encoder_forward(feats, w, shape)
    assert out.shape == (2, 1500, shape.d_model)
    pick = out[:, ::50, ::8].numpy()
    np.testing.assert_allclose(pick, gold[f"{shape_name}_out_pick"], atol=2e-4, rtol=0)
    np.testing.assert_allclose(pick, gold[f"{shape_name}_hf_pick"], atol=2e-4, rtol=0)
    probe, toks, margin = GreedyProbe.pick(out, shape.d_model, shape.n_heads)
    assert margin >= 0.1
    assert np.array_equal(toks.numpy(), gold[f"{shape_name}_probe_tokens"])


def test_invalid_feature_shapes_raise_value_error():
    shape = synth.SHAPES["micro"]
    w = synth.encoder_weights(shape)
    with pytest.raises(ValueError, match="Invalid input features shape"):
        encoder.encoder_forward(np.zeros((1, shape.n_mels + 1, 3000), np.float32), w, shape)
    with pytest.raises(ValueError, match="Invalid input features shape"):
        encoder.encoder_forward(np.zeros((1, shape.n_mels, 3001), np.float32), w, shape)


def test_flop_model_matches_survey_table():
    assert abs(synth.SHAPES["large-v3"].flops_per_window / 1e12 - 2.2738) < 1e-3
    assert abs(synth.SHAPES["medium"].flops_per_window / 1e12 - 1.1381) < 1e-3
    assert abs(synth.SHAPES["tiny"].flops_per_window / 1e12 - 0.0369) < 1e-3


def test_key_bias_is_absent():
    w = synth.encoder_weights(synth.SHAPES["micro"])
    d = synth.SHAPES["micro"].d_model
    assert not w["encoder/layer_0/self_attention/linear_0/bias"][d:2 * d].any()


def test_sinusoids_match_openai_definition():
    s = synth.sinusoids(1500, 384)
    assert s.shape == (1500, 384)
    assert np.allclose(s[0, :192], 0) and np.allclose(s[0, 192:], 1)
    assert abs(s[1, 0] - np.sin(1.0)) < 1e-6
