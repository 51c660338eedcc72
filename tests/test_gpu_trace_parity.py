"""GPU parity, part 1: rays and hits.  The sm_100a traversal (csrc/trace.cuh through the C ABI) must report the
same (instance, primitive, t, beta, gamma) -- bit for bit -- as the scalar oracle, for primary rays of the rtigo3
scenes and for incoherent rays, and the same visibility for shadow-type rays."""
import numpy as np
import pytest

import helpers as H
from oracle import orc
from tweeker_raytracer_b200 import core, host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cornell(cuda_device, tmp_path_factory):
    d = tmp_path_factory.mktemp("cornell")
    app = host.App(H.write_system(d, "rtigo3_cornell_box", resolution="256 256"), H.scene_path("rtigo3_cornell_box"))
    ref = H.oracle_scene(app)
    yield app, ref
    app.close()


@pytest.fixture(scope="module")
def geometry(cuda_device, tmp_path_factory):
    d = tmp_path_factory.mktemp("geometry")
    app = host.App(H.write_system(d, "rtigo3_geometry", resolution="320 180"), H.scene_path("rtigo3_geometry"))
    ref = H.oracle_scene(app)
    yield app, ref
    app.close()


def gpu_primary(app, iteration):
    ctx = app.context(0)
    sys = app.system_data(0)
    w, h = app.resolution
    d_rays = ctx.malloc(w * h * 32)
    ctx.generate_primary(sys, w, h, iteration, d_rays)
    rays = ctx.download(d_rays, core.RAY_DTYPE, w * h)
    ctx.free(d_rays)
    return rays


@pytest.mark.parametrize("iteration", [0, 7])
def test_primary_rays_bit_exact(cornell, iteration):
    app, ref = cornell
    w, h = app.resolution
    got = gpu_primary(app, iteration)
    want = ref.generate_primary(H.oracle_sys(app), w, h, iteration)
    assert got.tobytes() == want.tobytes()


def test_instance_inverse_matches_oracle(cornell):
    app, ref = cornell
    ctx = app.context(0)
    top = app.system_data(0).topObject
    for i in range(app.info.numInstances):
        assert ctx.instance_inverse(top, i).tobytes() == ref.inverse(i).tobytes()


@pytest.mark.parametrize("which", ["cornell", "geometry"])
def test_closest_hit_primary_bit_exact(which, cornell, geometry):
    app, ref = cornell if which == "cornell" else geometry
    ctx = app.context(0)
    top = app.system_data(0).topObject
    w, h = app.resolution
    rays = ref.generate_primary(H.oracle_sys(app), w, h, 3)
    got = ctx.trace_closest_host(top, rays)
    want = ref.trace_closest(rays)
    assert np.array_equal(got["inst"], want["inst"])
    assert np.array_equal(got["prim"], want["prim"])
    assert H.hits_equal(got, want)
    assert (want["inst"] != 0xffffffff).mean() > 0.5     # the scene fills most of the frame


@pytest.mark.parametrize("which", ["cornell", "geometry"])
def test_closest_hit_incoherent_bit_exact(which, cornell, geometry):
    app, ref = cornell if which == "cornell" else geometry
    ctx = app.context(0)
    top = app.system_data(0).topObject
    box = dict(lo=(-1.0, 0.0, -1.0), hi=(1.0, 2.0, 1.0)) if which == "cornell" else dict(lo=(-8, 0.01, -8), hi=(8, 5, 8))
    rays = H.random_rays(200000, seed=0x89ABCDEF, **box)
    got = ctx.trace_closest_host(top, rays)
    want = ref.trace_closest(rays)
    assert H.hits_equal(got, want)
    # ground truth on a subset: brute force over every triangle of every instance
    sub = rays[:300]
    assert H.hits_equal(got[:300], ref.trace_closest(sub, brute_force=True))


@pytest.mark.parametrize("which", ["cornell", "geometry"])
def test_any_hit_matches(which, cornell, geometry):
    app, ref = cornell if which == "cornell" else geometry
    ctx = app.context(0)
    top = app.system_data(0).topObject
    box = dict(lo=(-1.0, 0.0, -1.0), hi=(1.0, 2.0, 1.0)) if which == "cornell" else dict(lo=(-8, 0.01, -8), hi=(8, 5, 8))
    rays = H.random_rays(200000, seed=0x1234567, tmax=1.5, **box)
    got = ctx.trace_any_host(top, rays)
    want = ref.trace_any(rays)
    assert np.array_equal(got.astype(np.uint8), want)
    assert 0.05 < want.mean() < 0.999


def test_edge_cases(cornell):
    app, ref = cornell
    ctx = app.context(0)
    top = app.system_data(0).topObject
    rays = np.zeros(6, dtype=core.RAY_DTYPE)
    rays[0] = (0, 1, 0, 5e-5, 0, 0, -1, 1e27)          # axis-aligned direction (zero components)
    rays[1] = (0, 1, 0, 5e-5, 0, -1, 0, 0.5)           # tmax before the floor: miss
    rays[2] = (0, 1, 0, 2.0, 0, -1, 0, 1e27)           # tmin beyond the floor: miss
    rays[3] = (0, 1, 0, 1.0, 0, -1, 0, 1.0)            # empty interval
    rays[4] = (0, 1, 5, 5e-5, 0, 0, 1, 1e27)           # outside, pointing away
    rays[5] = (-1, 0, -1, 5e-5, 1, 1, 1, 1e27)         # starts on a corner, runs along the diagonal
    got = ctx.trace_closest_host(top, rays)
    want = ref.trace_closest(rays, brute_force=True)
    assert H.hits_equal(got, want)
    assert got["inst"][1] == 0xffffffff and got["t"][1] == -1.0
    # zero rays is a no-op
    assert len(ctx.trace_closest_host(top, rays[:0])) == 0


def test_trace_counts_are_deterministic(cornell):
    app, ref = cornell
    ctx = app.context(0)
    top = app.system_data(0).topObject
    rays = H.random_rays(50000, seed=5, lo=(-1, 0, -1), hi=(1, 2, 1))
    d = ctx.to_device(rays)
    a = ctx.trace_count(top, d, len(rays))
    b = ctx.trace_count(top, d, len(rays))
    ctx.free(d)
    assert (a.nodes, a.tris, a.instances, a.rays) == (b.nodes, b.tris, b.instances, b.rays)
    assert a.rays == len(rays) and a.nodes > a.rays and a.tris > 0 and a.instances > 0
