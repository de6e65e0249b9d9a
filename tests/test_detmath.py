"""The numerical contract (include/wifi_detmath.h) on its own: every function against libm / float64 over the
argument ranges the PHY uses, and -- on the GPU -- bit equality between the device and the host evaluation of the same
operation sequences.  Oracle and CUDA library share this header, so an error in it would be common-mode and invisible
to every oracle-vs-GPU comparison; this file is what looks at it from outside."""
import numpy as np
import pytest

SINCOS, ATAN2, LOG, CMUL, CMULC, CDIV, BOX_MULLER, CMAC, CMACC, CMSUBC, NORM_ADD, NORM_SUB, PHILOX = range(13)


def _angles(rng, n):
    """sync_short rotates by freq * j with |freq| <= pi/16 and j < 43200 (8.5e3 rad); the equalizer's arguments are small."""
    return np.concatenate([rng.uniform(-9000, 9000, n), rng.uniform(-7, 7, n), rng.uniform(-1e-3, 1e-3, n // 4),
                           np.array([0.0, -0.0, np.pi / 4, np.pi / 2, np.pi, -np.pi, 2 * np.pi, 1e-20, 8482.3])]).astype(np.float32)


def _points(rng, n):
    mag = 10.0 ** rng.uniform(-6, 6, n)
    ph = rng.uniform(-np.pi, np.pi, n)
    y, x = (mag * np.sin(ph)).astype(np.float32), (mag * np.cos(ph)).astype(np.float32)
    edge = np.array([(0, 0), (0, 1), (0, -1), (1, 0), (-1, 0), (1, 1), (-1, -1), (1e-30, 1), (1, 1e-30), (-1e-30, -1), (3, -3e-8)], np.float32)
    return np.concatenate([y, edge[:, 0]]), np.concatenate([x, edge[:, 1]])


def test_sincos_against_libm(O):
    rng = np.random.default_rng(1)
    x = _angles(rng, 1_000_000)
    s, c = O.detmath(SINCOS, x)
    x64 = x.astype(np.float64)
    # the float argument itself is exact; the reduction constants keep |error| below 3e-7 up to 1e4 rad
    assert np.abs(s - np.sin(x64)).max() < 3e-7 and np.abs(c - np.cos(x64)).max() < 3e-7
    assert np.abs(s * s.astype(np.float64) + c * c.astype(np.float64) - 1).max() < 5e-7
    s0, c0 = O.detmath(SINCOS, np.zeros(1, np.float32))
    assert s0[0] == 0 and c0[0] == 1


def test_atan2_against_libm(O):
    rng = np.random.default_rng(2)
    y, x = _points(rng, 1_000_000)
    r, _ = O.detmath(ATAN2, y, x)
    ref = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    err = np.abs(r - ref)
    err = np.minimum(err, np.abs(err - 2 * np.pi))          # (-pi, pi] wraps at the negative real axis
    assert err.max() < 4e-7, err.max()
    assert O.detmath(ATAN2, np.zeros(1, np.float32), np.zeros(1, np.float32))[0][0] == 0     # atan2(0, 0) = 0 as libm
    nan = O.detmath(ATAN2, np.array([np.nan, 1], np.float32), np.array([1, np.nan], np.float32))[0]
    assert np.isnan(nan).all()


def test_log_against_libm(O):
    rng = np.random.default_rng(3)
    # Box-Muller feeds u in (0, 1); the general range is checked too
    x = np.concatenate([rng.uniform(2.0 ** -25, 1.0, 500_000), 10.0 ** rng.uniform(-30, 30, 500_000), [1.0, 0.5, 2.0, np.e]]).astype(np.float32)
    r, _ = O.detmath(LOG, x)
    ref = np.log(x.astype(np.float64))
    assert (np.abs(r - ref) / np.maximum(1.0, np.abs(ref))).max() < 2e-7
    assert O.detmath(LOG, np.ones(1, np.float32))[0][0] == 0


def test_complex_products_against_float64(O):
    rng = np.random.default_rng(4)
    n = 500_000
    a, b, c, d = [(rng.standard_normal(n) * 10.0 ** rng.uniform(-3, 3, n)).astype(np.float32) for _ in range(4)]
    x, y = a.astype(np.float64) + 1j * b, c.astype(np.float64) + 1j * d
    scale = np.abs(x) * np.abs(y)
    for fn, ref in ((CMUL, x * y), (CMULC, x * np.conj(y))):
        re, im = O.detmath(fn, a, b, c, d)
        assert (np.abs((re + 1j * im) - ref) / scale).max() < 1.5e-7, fn      # one fused rounding + one product rounding per part
    re, im = O.detmath(CDIV, a, b, c, d)
    ref = x / y
    assert (np.abs((re + 1j * im) - ref) / np.abs(ref)).max() < 4e-7
    # multiply-accumulate forms: acc +- product
    e, f = [rng.standard_normal(n).astype(np.float32) for _ in range(2)]
    acc = e.astype(np.float64) + 1j * f
    for fn, ref in ((CMAC, acc + x * y), (CMACC, acc + x * np.conj(y)), (CMSUBC, acc - x * np.conj(y))):
        re, im = O.detmath(fn, a, b, c, d, e, f)
        assert (np.abs((re + 1j * im) - ref) / (scale + np.abs(acc))).max() < 1.5e-7, fn
    for fn, sgn in ((NORM_ADD, 1.0), (NORM_SUB, -1.0)):
        r, _ = O.detmath(fn, a, b, None, None, e)
        ref = e.astype(np.float64) + sgn * np.abs(x) ** 2
        assert (np.abs(r - ref) / (np.abs(x) ** 2 + np.abs(e))).max() < 1.5e-7, fn
    # exact cases the kernels rely on: multiplying by 1, by j, dividing by a real power of two
    one = np.ones(n, np.float32)
    re, im = O.detmath(CMUL, a, b, one, np.zeros(n, np.float32))
    assert np.array_equal(re, a) and np.array_equal(im, b)
    re, im = O.detmath(CDIV, a, b, 2 * one, np.zeros(n, np.float32))
    assert np.array_equal(re, a / 2) and np.array_equal(im, b / 2)


def test_philox_known_answer_and_box_muller_statistics(O):
    # Random123 known answer: philox4x32-10, counter = key = all ones
    ones = np.array([0xffffffff], np.uint32).view(np.float32)
    w0, w1 = O.detmath(PHILOX, ones, ones, ones, ones)
    # wdm_selftest passes (c0, c1, 0, 0) as counter: check against an independent Python implementation instead
    def philox(c, k):
        c, k = list(c), list(k)
        for _ in range(10):
            p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xffffffff, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xffffffff]
            k = [(k[0] + 0x9E3779B9) & 0xffffffff, (k[1] + 0xBB67AE85) & 0xffffffff]
        return c
    # the implementation above reproduces the Random123 vector for counter = key = ones ...
    assert philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    # ... and the contract's Philox equals it on the counters the selftest can express
    r = philox([0xffffffff, 0xffffffff, 0, 0], [0xffffffff, 0xffffffff])
    assert int(w0.view(np.uint32)[0]) == r[0] ^ r[2] and int(w1.view(np.uint32)[0]) == r[1] ^ r[3]
    rng = np.random.default_rng(5)
    a = rng.integers(0, 2 ** 32, 400_000, dtype=np.uint32).view(np.float32)
    b = rng.integers(0, 2 ** 32, 400_000, dtype=np.uint32).view(np.float32)
    z0, z1 = O.detmath(BOX_MULLER, a, b)
    z = np.concatenate([z0, z1]).astype(np.float64)
    assert np.isfinite(z).all() and abs(z.mean()) < 5e-3 and abs(z.var() - 1) < 1e-2 and abs((z ** 4).mean() - 3) < 0.1
    assert abs(np.mean(z0.astype(np.float64) * z1)) < 5e-3


@pytest.mark.gpu
def test_device_evaluates_the_contract_bit_for_bit(O, W):
    """Same operation sequences, same bits: sm_100a FFMA/FMUL/FADD/MUFU-free division vs x86-64 FMA3."""
    h = W.Handle(max_samples=1 << 16)
    rng = np.random.default_rng(6)
    try:
        x = _angles(rng, 1_000_000)
        for got, want in zip(h.detmath(SINCOS, x), O.detmath(SINCOS, x)):
            assert np.array_equal(got, want)
        y, xx = _points(rng, 1_000_000)
        assert np.array_equal(h.detmath(ATAN2, y, xx)[0], O.detmath(ATAN2, y, xx)[0], equal_nan=True)
        lx = np.concatenate([rng.uniform(2.0 ** -25, 1.0, 500_000), 10.0 ** rng.uniform(-30, 30, 500_000)]).astype(np.float32)
        assert np.array_equal(h.detmath(LOG, lx)[0], O.detmath(LOG, lx)[0])
        n = 500_000
        a, b, c, d, e, f = [(rng.standard_normal(n) * 10.0 ** rng.uniform(-3, 3, n)).astype(np.float32) for _ in range(6)]
        for fn in (CMUL, CMULC, CDIV, CMAC, CMACC, CMSUBC, NORM_ADD, NORM_SUB):
            g, w = h.detmath(fn, a, b, c, d, e, f), O.detmath(fn, a, b, c, d, e, f)
            assert np.array_equal(g[0], w[0]) and np.array_equal(g[1], w[1]), fn
        u = [rng.integers(0, 2 ** 32, n, dtype=np.uint32).view(np.float32) for _ in range(4)]
        for fn in (BOX_MULLER, PHILOX):
            g, w = h.detmath(fn, *u), O.detmath(fn, *u)
            assert np.array_equal(g[0].view(np.uint32), w[0].view(np.uint32)) and np.array_equal(g[1].view(np.uint32), w[1].view(np.uint32)), fn
    finally:
        h.close()
