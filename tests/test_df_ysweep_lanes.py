"""The y sweep of df_xy_dpx (voxelpathtracer_b200/csrc/df_build.cu, DF_Y_STEP) runs its two DPX chains on the HIGH bytes of the 16-bit lanes:
the odd bytes of an x-word are used as loaded, the even bytes after `word << 8`, the low byte of every lane holds junk.  This is the
arithmetic of that macro restated on Python integers (VIADDMNMX.U16x2 = per-lane min(a + b, c) with a wrapping add, PRMT = byte select)
and checked against (a) the clean-lane form it replaced and (b) the definition of the sweep — min over y' of v[y'] + |y - y'|, capped at
254 (Core/Shaders/ManhattanDistanceY.comp).  The kernel itself is checked on the GPU (tests/test_gpu_parity.py)."""
import numpy as np


def viaddmin_u16x2(a, b, c):
    r = 0
    for sh in (0, 16):
        s = (((a >> sh) & 0xFFFF) + ((b >> sh) & 0xFFFF)) & 0xFFFF
        r |= min(s, (c >> sh) & 0xFFFF) << sh
    return r


def byte_perm(x, y, sel):
    src = [(x >> (8 * i)) & 0xFF for i in range(4)] + [(y >> (8 * i)) & 0xFF for i in range(4)]
    return sum(src[(sel >> (4 * i)) & 0xF] << (8 * i) for i in range(4))


def sweep(words, init, step):
    a = b = init
    out = list(words)
    for order in (range(len(out)), range(len(out) - 1, -1, -1)):  # forward, then backward over what the forward pass stored
        for k in order:
            a, b, out[k] = step(a, b, out[k])
    return out


def step_clean(lo, hi, w):  # -DVXPT_DF_Y_CLEAN_LANES: byte pairs (0,1) / (2,3) unpacked into clean lanes
    lo = viaddmin_u16x2(lo, 0x00010001, byte_perm(w, 0, 0x4140))
    hi = viaddmin_u16x2(hi, 0x00010001, byte_perm(w, 0, 0x4342))
    return lo, hi, byte_perm(lo, hi, 0x6420)


def step_high(ce, co, w):  # default: chains on the lanes' high bytes
    co = viaddmin_u16x2(co, 0x01000100, w)
    ce = viaddmin_u16x2(ce, 0x01000100, (w << 8) & 0xFFFFFFFF)
    return ce, co, byte_perm(ce, co, 0x7351)


def columns(rng, kind):
    if kind == 0:
        return rng.integers(0, 255, size=(128, 4))                                   # anything an x sweep can leave, 254 included
    if kind == 1:
        return np.where(rng.random((128, 4)) < 0.05, 0, 254)                         # sparse solids, "no solid seen" elsewhere
    return np.minimum(254, rng.integers(0, 40, size=(128, 4)) * rng.integers(0, 8, size=(128, 4)))


def test_high_byte_chains_equal_clean_lanes_and_the_definition():
    rng = np.random.default_rng(20261018)
    for trial in range(45):
        b = columns(rng, trial % 3)
        words = [int(r[0]) | int(r[1]) << 8 | int(r[2]) << 16 | int(r[3]) << 24 for r in b]
        got = sweep(words, 0xFE00FE00, step_high)
        assert got == sweep(words, 0x00FE00FE, step_clean), trial
        y = np.arange(128)
        for c in range(4):
            ref = np.minimum(254, (b[:, c][None, :] + np.abs(y[:, None] - y[None, :])).min(axis=1))
            assert [(g >> (8 * c)) & 0xFF for g in got] == ref.tolist(), (trial, c)


def test_the_add_never_leaves_its_lane():
    # the largest value a chain can hold is 254 in the high byte with any low byte: + 0x0100 stays below 2^16
    assert viaddmin_u16x2(0xFEFFFEFF, 0x01000100, 0xFFFFFFFF) == 0xFFFFFFFF
    assert viaddmin_u16x2(0xFEFFFEFF, 0x01000100, 0xFE00FE00) == 0xFE00FE00
