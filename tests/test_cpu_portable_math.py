"""include/rt_portable_math.h: the pinned transcendentals stay within a few ulp of libm on the ranges the path tracer
uses.  Exercised through the two oracle builds (pinned vs libm) with a tiny C harness compiled on the fly."""
import numpy as np
import pytest

import helpers as H

@pytest.fixture(scope="module")
def pm(tmp_path_factory):
    return H.portable_math_lib(tmp_path_factory.mktemp("pm"))


def run(lib, name, x, y=None):
    return H.portable_math_call(lib, name[2:], x, y)


def ulps(a, b):
    b = b.astype(np.float32)
    return np.abs(a.astype(np.float64) - b.astype(np.float64)) / np.maximum(np.spacing(np.abs(b)).astype(np.float64), 1e-45)


def test_sin_cos(pm):
    x = np.linspace(-4 * np.pi, 4 * np.pi, 200001)
    for name, f in (("p_sin", np.sin), ("p_cos", np.cos)):
        got = run(pm, name, x)
        want = f(x.astype(np.float32).astype(np.float64))
        assert np.max(np.abs(got - want)) < 2.5e-7


def test_atan_atan2_acos(pm):
    x = np.concatenate([np.linspace(-50, 50, 100001), [1e6, -1e6, 0.0]])
    assert ulps(run(pm, "p_atan", x), np.arctan(x.astype(np.float32).astype(np.float64))).max() <= 3
    rng = np.random.default_rng(1)
    a, b = rng.normal(size=100000), rng.normal(size=100000)
    assert np.max(np.abs(run(pm, "p_atan2", a, b) - np.arctan2(a.astype(np.float32).astype(np.float64), b.astype(np.float32).astype(np.float64)))) < 5e-7
    assert run(pm, "p_atan2", [0.0, 1.0, -1.0, 0.0], [0.0, 0.0, 0.0, -1.0]).tolist() == pytest.approx([0.0, np.pi / 2, -np.pi / 2, np.pi], abs=1e-7)
    x = np.linspace(-1, 1, 100001)
    assert np.max(np.abs(run(pm, "p_acos", x) - np.arccos(x.astype(np.float32).astype(np.float64)))) < 5e-7
    assert run(pm, "p_acos", [1.0000001, -1.0000001]).tolist() == pytest.approx([0.0, np.pi], abs=1e-6)


def test_exp_log_pow(pm):
    x = np.linspace(-80, 10, 100001)
    assert ulps(run(pm, "p_exp", x), np.exp(x.astype(np.float32).astype(np.float64))).max() <= 3
    assert run(pm, "p_exp", [0.0])[0] == 1.0 and run(pm, "p_exp", [-200.0])[0] == 0.0
    x = np.concatenate([np.logspace(-30, 30, 50001), [1.0]])
    assert ulps(run(pm, "p_log", x), np.log(x.astype(np.float32).astype(np.float64))).max() <= 3
    b = np.linspace(0, 4, 4001)
    for e in (1.0 / 2.2, 1.4, 2.2):
        got = run(pm, "p_pow", b, np.full_like(b, e))
        want = np.power(b.astype(np.float32).astype(np.float64), np.float32(e).astype(np.float64))
        assert np.max(np.abs(got - want) / np.maximum(want, 1e-3)) < 2e-6
    assert run(pm, "p_pow", [0.0, -1.0], [2.0, 2.0]).tolist() == [0.0, 0.0]


def test_ieee_division_sqrt_and_uncontracted_multiply_add(pm):
    # the three operations the shading kernels need compiler flags for (-prec-div, -prec-sqrt, -fmad=false) against float64
    rng = np.random.default_rng(5)
    a = rng.normal(size=200000).astype(np.float32) * np.float32(10.0) ** rng.integers(-10, 10, size=200000).astype(np.float32)
    b = rng.normal(size=200000).astype(np.float32) * np.float32(10.0) ** rng.integers(-10, 10, size=200000).astype(np.float32)
    b[b == 0] = 1.0
    assert np.array_equal(run(pm, "p_div", a, b), (a.astype(np.float64) / b.astype(np.float64)).astype(np.float32))
    assert np.array_equal(run(pm, "p_sqrt", np.abs(a)), np.sqrt(np.abs(a).astype(np.float64)).astype(np.float32))
    two_roundings = ((a.astype(np.float64) * b.astype(np.float64)).astype(np.float32).astype(np.float64) + a.astype(np.float64)).astype(np.float32)
    assert np.array_equal(run(pm, "p_muladd", a, b), two_roundings)
