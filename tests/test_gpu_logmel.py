"""GPU parity: the sm_100a log-mel kernel, through the C ABI, against the numpy oracle and the golden fixtures.
Tolerance (BASELINE.json north_star): 1e-4 absolute on the f32 log-mel."""
import os

import numpy as np
import pytest
import torch

from oracle import logmel, synth

pytestmark = pytest.mark.gpu
TOL = 1e-4
GENS = {"tone": (synth.tone_noise, 0), "chirp": (synth.am_chirp, 1), "gapped": (synth.gapped, 2)}


@pytest.fixture(scope="module")
def fes():
    from whisper_aries_b200 import FeatureExtractor
    return {n: FeatureExtractor(feature_size=n) for n in (80, 128)}


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "logmel_golden.npz"))


@pytest.mark.parametrize("n_mels", [80, 128])
@pytest.mark.parametrize("name", list(GENS))
def test_full_window_vs_oracle_and_golden(fes, gold, n_mels, name):
    gen, seed = GENS[name]
    x = gen(seed)
    got = fes[n_mels](x)
    assert got.shape == (n_mels, 3001) and got.dtype == np.float32
    assert np.abs(got - logmel.log_mel(x, n_mels)).max() <= TOL
    pick = gold["frame_pick"]
    key = f"m{n_mels}_{name}"
    assert np.abs(got[:, pick] - gold[key + "_pick"]).max() <= TOL               # committed oracle output
    hf = pick[pick < 2998]
    assert np.abs(got[:, hf] - gold[key + "_hf_pick"]).max() <= 2e-4             # independent implementation (HF)


@pytest.mark.parametrize("n_mels", [80, 128])
@pytest.mark.parametrize("n", [16000, 8000, 1234, 29600])
def test_short_and_ragged_vs_golden(fes, gold, n_mels, n):
    got = fes[n_mels](synth.window_signal(7, n))
    assert got.shape == (n_mels, (n + 160) // 160)
    assert np.abs(got - gold[f"m{n_mels}_short{n}"]).max() <= TOL


@pytest.mark.parametrize("n", [1, 41, 159, 160, 201, 10240 + 37, 185 * 16000])
def test_edge_lengths(fes, n):
    x = synth.window_signal(11, n)          # 185 s is the reference's own work-item length (final_optimized_transcriber.py:422-449)
    ref = logmel.log_mel(x, 80)
    got = fes[80](x)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= TOL


def test_silence_padding_and_clamp(fes):
    fe = fes[80]
    assert np.allclose(fe(np.zeros(480000, np.float32)), -1.5, atol=1e-6)
    x = synth.tone_noise(3, 32000)
    x[16000:] = 0.0
    x[100] = 50.0
    got = fe(x)
    assert np.abs(got - logmel.log_mel(x, 80)).max() <= TOL
    assert abs(got.min() - (got.max() - 2.0)) < 1e-5                             # floor = max - 8 decades
    assert np.abs(fe(x, padding=0) - logmel.log_mel(x, 80, padding=0)).max() <= TOL
    assert fe(np.float64(x)).dtype == np.float32                                  # upstream casts its input


def test_batch_device_and_window_layout(fes):
    fe = fes[128]
    xs = synth.batch_signals(7, 20)
    ref = np.stack([logmel.log_mel_window(x, 128) for x in xs])
    host = fe(xs, frames_out=3000)
    assert host.shape == (7, 128, 3000) and np.abs(host - ref).max() <= TOL
    devt = fe(torch.from_numpy(xs).cuda(), frames_out=3000)
    assert devt.is_cuda and np.array_equal(devt.cpu().numpy(), host)             # same kernel, same bits
    # strided batch (rows of a bigger buffer) and a short signal padded to the 3000-frame window
    big = torch.zeros((3, 500000), device="cuda")
    big[:, :480000] = torch.from_numpy(xs[:3]).cuda()
    assert np.array_equal(fe(big[:, :480000], frames_out=3000).cpu().numpy(), host[:3])
    short = fe(xs[0, :16000], frames_out=3000)
    assert np.abs(short - logmel.pad_or_trim(logmel.log_mel(xs[0, :16000], 128))).max() <= TOL
    assert (short[:, 101:] == 0).all()


def test_full_size_properties(fes):
    """BASELINE config 3 (64 windows): size-independent checks — per-item independence, shift by one hop, scale law."""
    fe = fes[128]
    xs = torch.from_numpy(synth.batch_signals(6, 0)).cuda().repeat(11, 1)[:64].contiguous()
    out = fe(xs, frames_out=3000)
    assert out.shape == (64, 128, 3000) and torch.isfinite(out).all()
    assert torch.equal(out[0], out[6]) and torch.equal(out[5], out[59])           # items do not leak into each other
    one = fe(xs[13:14], frames_out=3000)
    assert torch.equal(one[0], out[13])                                           # batch size does not change bits
    # scaling the PCM by 2^k shifts every unclamped value by k * log10(4) / 4 exactly in exact arithmetic
    a = fe(xs[0], frames_out=3000)
    b = fe(xs[0] * 4.0, frames_out=3000)
    assert (b - a - np.log10(16.0) / 4).abs().max().item() <= 2e-5
    assert fe.last_launches == 2
