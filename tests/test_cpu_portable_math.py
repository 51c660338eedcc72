"""include/rt_portable_math.h: the pinned transcendentals stay within a few ulp of libm on the ranges the path tracer
uses.  Exercised through the two oracle builds (pinned vs libm) with a tiny C harness compiled on the fly."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import helpers as H

SRC = r"""
#include "rt_portable_math.h"
#define W(name, expr) void name(const float* x, const float* y, float* o, int n) { for (int i = 0; i < n; ++i) o[i] = expr; }
W(p_sin, rt_sinf(x[i])) W(p_cos, rt_cosf(x[i])) W(p_atan, rt_atanf(x[i])) W(p_atan2, rt_atan2f(x[i], y[i]))
W(p_acos, rt_acosf(x[i])) W(p_exp, rt_expf(x[i])) W(p_log, rt_logf(x[i])) W(p_pow, rt_powf(x[i], y[i]))
"""


@pytest.fixture(scope="module")
def pm(tmp_path_factory):
    d = tmp_path_factory.mktemp("pm")
    src = d / "pm.c"
    src.write_text(SRC)
    so = d / "pm.so"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I" + os.path.join(H.ROOT, "include"), "-o", str(so), str(src), "-lm"])
    return C.CDLL(str(so))


def run(lib, name, x, y=None):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y if y is not None else np.zeros_like(x), dtype=np.float32)
    o = np.zeros_like(x)
    getattr(lib, name)(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), len(x))
    return o


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
