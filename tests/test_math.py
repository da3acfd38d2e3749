"""The shared deterministic math header (csrc/ktn_math.h) against mpmath."""
import ctypes as C

import mpmath as mp
import numpy as np
import pytest

mp.mp.prec = 200


def _call(lib, fn, *arrs):
    n = len(arrs[0]); out = np.empty(n)
    getattr(lib.dll, fn)(*[a.ctypes.data_as(C.c_void_p) for a in arrs], out.ctypes.data_as(C.c_void_p), C.c_int64(n))
    return out


def _max_ulp(got, exact):
    worst = 0.0
    for g, e in zip(got, exact):
        ef = float(e)
        if ef == 0 or not np.isfinite(ef):
            continue
        worst = max(worst, float(abs(mp.mpf(float(g)) - e) / mp.mpf(float(np.spacing(abs(ef))))))
    return worst


def test_exp_below_one_ulp(emu_lib):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-5, 5, 1500), rng.uniform(-708, 708, 1000), rng.uniform(-1e-3, 1e-3, 200), rng.uniform(-745, -708, 200), rng.uniform(708, 709.7, 100)])
    assert _max_ulp(_call(emu_lib, "ktn_test_exp", x), [mp.exp(mp.mpf(float(v))) for v in x]) < 1.0


def test_log_below_one_ulp(emu_lib):
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(0.5, 2, 1500), 10.0**rng.uniform(-300, 300, 1000), rng.uniform(0.99, 1.01, 500), 10.0**rng.uniform(-320, -308, 100)])
    assert _max_ulp(_call(emu_lib, "ktn_test_log", x), [mp.log(mp.mpf(float(v))) for v in x]) < 1.0


def test_sin_cos_below_one_ulp(emu_lib):
    rng = np.random.default_rng(4)
    near = np.array([k * np.pi / 2 for k in range(-40, 41)]) * (1 + rng.uniform(-1e-15, 1e-15, 81))      # doubles next to multiples of pi/2
    x = np.concatenate([rng.uniform(-np.pi, np.pi, 2000), rng.uniform(-100, 100, 1500), rng.uniform(-1.6e6, 1.6e6, 1500), rng.uniform(-1e-4, 1e-4, 200),
                        near, np.array([np.pi / 4, -np.pi / 4, 3 * np.pi / 4, 1e-300, 5e-324, 355.0, 1647099.0])])
    assert _max_ulp(_call(emu_lib, "ktn_test_sin", x), [mp.sin(mp.mpf(float(v))) for v in x]) < 1.0
    assert _max_ulp(_call(emu_lib, "ktn_test_cos", x), [mp.cos(mp.mpf(float(v))) for v in x]) < 1.0
    big = rng.uniform(-1e9, 1e9, 500)                       # beyond 2^20 * pi/2: the error grows with |x| but stays far inside 1e-12 relative to 1
    for fn, ref in (("ktn_test_sin", mp.sin), ("ktn_test_cos", mp.cos)):
        got = _call(emu_lib, fn, big)
        assert max(abs(float(mp.mpf(float(g)) - ref(mp.mpf(float(v))))) for g, v in zip(got, big)) < 1e-15


def test_pow_accuracy(emu_lib):
    rng = np.random.default_rng(3)
    x = 10.0**rng.uniform(-3, 3, 2000); p = rng.uniform(-8, 8, 2000)
    assert _max_ulp(_call(emu_lib, "ktn_test_pow", x, p), [mp.power(mp.mpf(float(a)), mp.mpf(float(b))) for a, b in zip(x, p)]) < 2.0
    x = np.full(500, np.e); p = rng.uniform(-20, 20, 500)     # e^(x) as the reference's tests write it
    assert _max_ulp(_call(emu_lib, "ktn_test_pow", x, p), [mp.power(mp.mpf(float(a)), mp.mpf(float(b))) for a, b in zip(x, p)]) < 2.0


def test_special_values(emu_lib):
    inf, nan = np.inf, np.nan
    e = _call(emu_lib, "ktn_test_exp", np.array([inf, -inf, nan, 710.0, -746.0, 0.0]))
    assert e[0] == inf and e[1] == 0 and np.isnan(e[2]) and e[3] == inf and e[4] == 0 and e[5] == 1
    l = _call(emu_lib, "ktn_test_log", np.array([inf, -1.0, nan, 0.0, 1.0, 5e-324]))
    assert l[0] == inf and np.isnan(l[1]) and np.isnan(l[2]) and l[3] == -inf and l[4] == 0 and abs(l[5] + 744.4400719213812) < 1e-12
    sn = _call(emu_lib, "ktn_test_sin", np.array([0.0, -0.0, inf, -inf, nan, 2.0**45, 1e300]))
    assert sn[0] == 0 and not np.signbit(sn[0]) and sn[1] == 0 and np.signbit(sn[1]) and all(np.isnan(sn[2:]))
    cs = _call(emu_lib, "ktn_test_cos", np.array([0.0, -0.0, inf, -inf, nan, 2.0**45, 1e300]))
    assert cs[0] == 1 and cs[1] == 1 and all(np.isnan(cs[2:]))
    x = np.array([0.0, -0.0, inf, -inf, -2.0, -2.0, -2.0, 0.0, 2.0, 0.5, 1.0, -8.0, 3.0])
    p = np.array([2.0, 3.0, 2.0, 3.0, 2.0, 3.0, 0.5, -1.0, inf, inf, nan, 1 / 3, 0.0])
    got = _call(emu_lib, "ktn_test_pow", x, p)
    want = [0.0, -0.0, inf, -inf, 4.0, -8.0, nan, inf, inf, 0.0, 1.0, nan, 1.0]
    for g, w in zip(got, want):
        assert (np.isnan(g) and np.isnan(w)) or (g == w and np.signbit(g) == np.signbit(w))
