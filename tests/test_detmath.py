"""include/phd_detmath.h: the shared deterministic fp32 functions are within a few ulp of float64 libm
over the ranges the filter uses, and Philox4x32-10 matches its published known-answer vectors."""
import numpy as np

from oracle import oracle as O


def ulp_err(got, ref64):
    ref32 = ref64.astype(np.float32)
    ulp = np.spacing(np.abs(ref32)).astype(np.float64)
    return np.abs(got.astype(np.float64) - ref64) / np.maximum(ulp, np.finfo(np.float32).tiny)


def test_exp():
    x = np.concatenate([np.linspace(-87, 88, 200001), np.random.default_rng(0).uniform(-30, 5, 100000)]).astype(np.float32)
    got = O.detmath("exp", x)
    assert ulp_err(got, np.exp(x.astype(np.float64))).max() < 2.0
    assert O.detmath("exp", np.float32([-88, -1e30, -3.4e38]))[0:3].tolist() == [0, 0, 0]
    assert np.isinf(O.detmath("exp", np.float32([89.0]))[0])


def test_log():
    rng = np.random.default_rng(1)
    x = np.concatenate([np.exp(rng.uniform(-80, 80, 200000)), rng.uniform(0.5, 2, 100000), [1.0, 1e-40]]).astype(np.float32)
    got = O.detmath("log", x)
    ref = np.log(x.astype(np.float64))
    err = np.abs(got.astype(np.float64) - ref) / np.maximum(np.spacing(np.abs(ref.astype(np.float32))), 1e-7)
    assert err.max() < 3.0
    assert O.detmath("log", np.float32([1.0]))[0] == 0.0
    # safeLog (reference src/device_math.cuh:9-16): x <= 0 -> -FLT_MAX
    assert O.detmath("safe_log", np.float32([0.0, -1.0])).tolist() == [-np.finfo(np.float32).max] * 2


def test_atan2_sincos_tan_wrap():
    rng = np.random.default_rng(2)
    x = rng.uniform(-20, 20, 200000).astype(np.float32)
    y = rng.uniform(-20, 20, 200000).astype(np.float32)
    got = O.detmath("atan2", x, y)
    ref = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    assert np.abs(got - ref).max() < 5e-7
    th = rng.uniform(-50, 50, 200000).astype(np.float32)
    s, c = O.detmath("sincos", th)
    assert np.abs(s - np.sin(th.astype(np.float64))).max() < 2.5e-7
    assert np.abs(c - np.cos(th.astype(np.float64))).max() < 2.5e-7
    a = rng.uniform(-1.2, 1.2, 100000).astype(np.float32)
    t = O.detmath("tan", a)
    assert (np.abs(t - np.tan(a.astype(np.float64))) / np.maximum(1, np.abs(t))).max() < 5e-7
    # wrapAngle (src/device_math.cuh:242-251) restated with double arithmetic
    w = np.concatenate([rng.uniform(-13, 13, 200000), rng.uniform(-1000, 1000, 1000)]).astype(np.float32)
    got = O.detmath("wrap", w)
    r = np.fmod(w.astype(np.float64), np.float64(np.float32(2 * np.pi)))
    r = np.where(r > np.pi, r - 2 * np.pi, np.where(r < -np.pi, r + 2 * np.pi, r))
    assert np.abs(got - r).max() < 3e-7
    assert np.abs(got).max() <= np.float32(np.pi)


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32-10
    lib = O.load()
    out = (np.zeros(4, dtype=np.uint32))
    lib.oracle_philox(0, 0, 0, 0, 0, 0, out.ctypes.data)
    assert [hex(v) for v in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    lib.oracle_philox(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, out.ctypes.data)
    assert [hex(v) for v in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    lib.oracle_philox(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0, out.ctypes.data)
    assert [hex(v) for v in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_warp_sum_shape():
    lib = O.load()
    v = np.random.default_rng(3).uniform(0, 1, 1000).astype(np.float32)
    got = lib.oracle_warp_sum(v.ctypes.data, len(v))
    # same tree in numpy: 64 strided partials, adjacent pairs added, xor butterfly over 32 lanes
    p = np.zeros(64, dtype=np.float32)
    for l in range(64):
        acc = np.float32(0)
        for x in v[l::64]:
            acc = np.float32(acc + x)
        p[l] = acc
    p = (p[0::2] + p[1::2]).astype(np.float32)
    for off in (16, 8, 4, 2, 1):
        p = (p + p[np.arange(32) ^ off]).astype(np.float32)
    assert got == p[0]
    assert abs(got - v.astype(np.float64).sum()) < 1e-3
