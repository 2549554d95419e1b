"""The denoiser's pinned transcendentals (voxelpathtracer_b200/csrc/denoise.cu: exp_cr, pow01_cr / pow_lt1_cr, normal_weight) are short
double-precision evaluations with a rounding test that falls back to the library function.  The contract is that they return, for EVERY
input, the value of the expression they replace — (float)exp((double)x), (float)pow((double)x, (double)y), pow(max(dot(n_a, n_b), floor), y)
over the face-normal ids — so the kernels' source is compiled for the host (tests/host_shadow) and compared with those expressions, evaluated
in numpy's double precision (glibc: correctly rounded in all but a vanishing share of cases, which the rounding test routes to the library
anyway), over tens of millions of inputs."""
import ctypes as C

import numpy as np
import pytest

from host_shadow import kernels_on_host as koh

pytestmark = pytest.mark.skipif(not koh.available(), reason="CUDA toolkit headers not present")


def _exp_cr(x):
    y = np.empty_like(x)
    koh.load().hs_exp_cr(x.ctypes.data, y.ctypes.data, x.size)
    return y


def _pow01_cr(x, e):
    y = np.empty_like(x)
    koh.load().hs_pow01_cr(x.ctypes.data, e.ctypes.data, y.ctypes.data, x.size)
    return y


def _same(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


def test_exp_equals_the_double_precision_exponential_rounded_once():
    # every 53rd float of [-104.5, -0): 21 million arguments, dense in every binade the filters can produce
    bits = np.arange(0x80000001, 0xC2D10000, 53, dtype=np.uint64).astype(np.uint32)
    x = bits.view(np.float32)
    assert x.min() < -104.0 and x.max() < 0
    with np.errstate(under="ignore"):
        want = np.exp(x.astype(np.float64)).astype(np.float32)
    got = _exp_cr(x)
    bad = ~_same(got, want)
    assert not bad.any(), (x[bad][:5], got[bad][:5], want[bad][:5])
    # the short evaluation carries nearly all of them: its rounding test refuses about 33 / 2^29 of the inputs
    inside = np.ascontiguousarray(x[(x > -87.0)])
    refused = koh.load().hs_exp_short_refused(inside.ctypes.data, inside.size)
    assert refused < 1e-5 * inside.size, refused


def test_exp_special_values():
    x = np.array([0.0, -0.0, -np.inf, np.inf, np.nan, 1.0, 88.0, 89.0, -87.0, -87.5, -103.9, -104.0, -1e-30, -1e-45, 1e-20, -150.0, -3e38], np.float32)
    with np.errstate(all="ignore"):
        want = np.exp(x.astype(np.float64)).astype(np.float32)
    assert _same(_exp_cr(x), want).all()


def test_pow_equals_the_double_precision_power_rounded_once():
    rng = np.random.default_rng(11)
    n = 6_000_000
    # the shapes the filters produce: a base in (0, 1] (1 - |difference| / 3, 1 - variance, e^-d) and an exponent of 0.1 .. 134
    x = np.concatenate([rng.random(n, dtype=np.float32), 1.0 - rng.random(n, dtype=np.float32) * np.float32(0.1),
                        np.exp(-rng.random(n, dtype=np.float32) * 4).astype(np.float32)])
    e = np.concatenate([rng.uniform(0.05, 24.0, n).astype(np.float32), rng.choice(np.array([76, 90, 102, 118, 134], np.float32), n),
                        np.full(n, 48.0, np.float32)])
    with np.errstate(under="ignore"):
        want = np.power(x.astype(np.float64), e.astype(np.float64)).astype(np.float32)
    got = _pow01_cr(x, e)
    bad = ~_same(got, want)
    assert not bad.any(), (x[bad][:5], e[bad][:5], got[bad][:5], want[bad][:5])
    assert (got > 0).mean() > 0.5       # not a test of underflow to zero


def test_pow_special_values():
    x = np.array([0.0, 1.0, 0.5, 0.5, 1e-9, 3.0, 0.999, np.nan, 0.5, 2.0, 1e-30, 0.25, -0.5], np.float32)
    e = np.array([3.0, 7.0, 1e7, 150.0, 32.0, 32.0, 1e-8, 2.0, np.nan, 0.5, 4.0, 0.5, 2.0], np.float32)
    with np.errstate(all="ignore"):
        want = np.power(x.astype(np.float64), e.astype(np.float64)).astype(np.float32)
    assert _same(_pow01_cr(x, e), want).all()


def test_normal_weight_equals_the_power_of_the_clamped_dot_product():
    def normal(i):
        return {0: (0, 0, 1), 1: (0, 0, -1), 2: (0, 1, 0), 3: (0, -1, 0), 4: (-1, 0, 0), 5: (1, 0, 0)}.get(i, (1, 1, 1))
    lib = koh.load()
    for a in list(range(9)) + [127, 255]:
        for b in list(range(9)) + [127, 255]:
            d = np.float32(sum(np.float32(p) * np.float32(q) for p, q in zip(normal(a), normal(b))))
            for floor, power in ((0.0, 16), (0.0, 32), (1e-9, 32)):
                want = np.float32(np.power(np.float64(max(d, np.float32(floor))), float(power)))
                got = lib.hs_normal_weight(a, b, C.c_float(0.0), power)
                assert np.float32(got) == want, (a, b, floor, power, got, want)
