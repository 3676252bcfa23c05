"""CPU check of the polynomial GELU used by the GEMM epilogues (ptx.cuh: gelu_poly2): the same f32 Horner evaluation,
restated in numpy, against the exact-erf definition the reference uses (CT2 ops::GELU / torch.nn.GELU)."""
import math
import re

import numpy as np

COEFFS_RE = re.compile(r"make_float2\((-?[0-9.e+-]+)f, ")


def coeffs_from_source():
    import os
    src = open(os.path.join(os.path.dirname(__file__), "..", "whisper_aries_b200", "csrc", "ptx.cuh")).read()
    body = src[src.index("gelu_poly2(float2 x)"):]
    body = body[:body.index("return __fmul2_rn")]
    body = body[body.index("const float2 u ="):]
    vals = [np.float32(v) for v in COEFFS_RE.findall(body)]
    return vals[:-1], vals[-1]          # Horner coefficients high -> low, then the 0.5


def gelu_poly(x):
    cs, half = coeffs_from_source()
    x = x.astype(np.float32)
    xc = np.clip(x, np.float32(-4.25), np.float32(4.25))
    u = (xc * xc).astype(np.float32)
    q = np.full_like(u, cs[0])
    for c in cs[1:]:
        q = (q * u + c).astype(np.float32)
    xm = np.maximum(x, np.float32(-4.25))
    return (xm * (xc * q + half).astype(np.float32)).astype(np.float32)


def test_polynomial_gelu_matches_erf_gelu():
    cs, half = coeffs_from_source()
    assert len(cs) == 9 and half == np.float32(0.5)
    x = np.concatenate([np.linspace(-12, 12, 600001), np.array([0.0, -0.0, 1e-20, -1e-20, 50.0, -50.0, 1e4, -1e4])])
    ref = np.array([0.5 * v * (1.0 + math.erf(v / math.sqrt(2.0))) for v in x])
    err = np.abs(gelu_poly(x).astype(np.float64) - ref)
    rel = err / np.maximum(1.0, np.abs(ref))
    assert rel.max() <= 9.5e-5, (rel.max(), x[rel.argmax()])
    assert err[np.abs(x) <= 3].max() <= 4e-5
    big = gelu_poly(np.array([1e4, -1e4]))
    assert abs(big[0] - 1e4) <= 0.2 and abs(big[1]) <= 1e-4
