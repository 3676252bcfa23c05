"""CPU: the log-mel kernel's per-thread code (whisper_aries_b200/csrc/logmel_core.cuh) walked thread by thread in plain
C++ (tests/emu) against the numpy oracle — checks the FFT factorisation, frame pairing, reflect / zero padding, tiling
and two-pass clamp without a GPU.  Tolerance is the path's stated one: 1e-4 absolute."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import logmel, synth

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "emu", "liblogmel_emu.so")
FP = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "emu", "logmel_emu.cpp")
    core = os.path.join(HERE, "..", "whisper_aries_b200", "csrc", "logmel_core.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", SO, src])
    lib = ctypes.CDLL(SO)
    lib.emu_logmel.argtypes = [FP, ctypes.c_longlong, ctypes.c_int, FP, ctypes.c_int, FP, ctypes.c_int]
    return lib


def run(lib, x, n_mels, padding=160, frames_out=None):
    filt = logmel.mel_filterbank(n_mels)
    n_frames = (len(x) + padding) // 160
    frames_out = n_frames if frames_out is None else frames_out
    out = np.full((n_mels, frames_out), np.nan, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    r = lib.emu_logmel(x.ctypes.data_as(FP), len(x), padding, filt.ctypes.data_as(FP), n_mels, out.ctypes.data_as(FP),
                       frames_out)
    assert r == n_frames
    return out


def test_fft20_matches_numpy(emu):
    rng = np.random.default_rng(0)
    for _ in range(5):
        x = (rng.standard_normal(20) + 1j * rng.standard_normal(20)).astype(np.complex64)
        ir, ii = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
        orr, oi = np.zeros(20, np.float32), np.zeros(20, np.float32)
        emu.emu_fft20(ir.ctypes.data_as(FP), ii.ctypes.data_as(FP), orr.ctypes.data_as(FP), oi.ctypes.data_as(FP))
        assert np.abs((orr + 1j * oi) - np.fft.fft(x.astype(np.complex128))).max() < 5e-6


@pytest.mark.parametrize("n_mels", [80, 128])
@pytest.mark.parametrize("name", ["tone", "chirp", "gapped"])
def test_full_window(emu, n_mels, name):
    gen, seed = {"tone": (synth.tone_noise, 0), "chirp": (synth.am_chirp, 1), "gapped": (synth.gapped, 2)}[name]
    x = gen(seed)
    assert np.abs(run(emu, x, n_mels) - logmel.log_mel(x, n_mels)).max() <= 1e-4


@pytest.mark.parametrize("n", [1, 41, 160, 199, 1234, 8000, 10240 + 37, 29600])
def test_short_and_ragged(emu, n):
    x = synth.window_signal(7, n)
    ref = logmel.log_mel(x, 80)
    got = run(emu, x, 80)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-4


def test_padding_zero_and_pad_or_trim(emu):
    x = synth.window_signal(3, 16000)
    assert np.abs(run(emu, x, 80, padding=0) - logmel.log_mel(x, 80, padding=0)).max() <= 1e-4
    got = run(emu, x, 80, frames_out=3000)                   # zero-filled past the last real frame (row a-4)
    ref = logmel.pad_or_trim(logmel.log_mel(x, 80))
    assert np.abs(got - ref).max() <= 1e-4 and (got[:, 101:] == 0).all()
    full = synth.tone_noise(0)
    got = run(emu, full, 128, frames_out=3000)               # trimmed: frame 3000 still counts for the clamp maximum
    assert np.abs(got - logmel.log_mel_window(full, 128)).max() <= 1e-4


def test_silence_and_clamp(emu):
    assert np.allclose(run(emu, np.zeros(48000, np.float32), 80), -1.5, atol=1e-6)
    x = synth.tone_noise(3, 32000)
    x[16000:] = 0.0
    x[100] = 50.0                                            # a click far from the silent half lifts its floor
    assert np.abs(run(emu, x, 80) - logmel.log_mel(x, 80)).max() <= 1e-4
