"""GPU parity, part 3: the device LBVH builder.  Closest hits are defined independently of the acceleration structure
(smallest t, ties -> lowest (instance, primitive)), so scenes built by the device builder must give bit-identical hits
and frames to the oracle, and to the host SAH builder."""
import os

import numpy as np
import pytest

import helpers as H
from oracle import orc
from tweeker_raytracer_b200 import core, host

pytestmark = pytest.mark.gpu


def soup(n, seed, extent=0.05):
    """Triangle soup: centres uniform in the unit cube, edge length ~extent, as TriangleAttributes + indices."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(0, 1, size=(n, 1, 3))
    v = (c + rng.uniform(-extent, extent, size=(n, 3, 3))).astype(np.float32).reshape(-1, 3)
    attrs = np.zeros(3 * n, dtype=orc.ATTR_DTYPE)
    attrs["vertex"] = v
    attrs["normal"] = (0, 0, 1)
    attrs["tangent"] = (1, 0, 0)
    return attrs, np.arange(3 * n, dtype=np.uint32).reshape(n, 3)


def build_scene(ctx, attrs, idx, transforms, flags):
    d_a, d_i = ctx.to_device(attrs), ctx.to_device(idx)
    gas = ctx.gas_build(d_a, 48, len(attrs), d_i, len(idx), flags)
    inst = np.zeros(len(transforms), dtype=core.INSTANCE_DTYPE)
    for k, t in enumerate(transforms):
        inst[k]["transform"], inst[k]["instanceId"], inst[k]["gas"], inst[k]["materialIndex"], inst[k]["lightIndex"] = t, k, gas, 0, -1
    return ctx.ias_build(inst), (d_a, d_i)


def oracle_for(attrs, idx, transforms):
    s = orc.Scene()
    g = s.add_geometry(attrs, idx)
    for t in transforms:
        s.add_instance(t, g, 0, -1)
    s.set_materials(np.zeros(1, dtype=orc.MATERIAL_DTYPE))
    s.set_lights(np.zeros(0, dtype=orc.LIGHT_DTYPE))
    s.set_camera(np.zeros(1, dtype=orc.CAMERA_DTYPE))
    s.commit()
    return s


IDENTITY = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]


@pytest.mark.parametrize("n,extent", [(1, 0.3), (2, 0.3), (7, 0.2), (1000, 0.08), (200000, 0.01)])
def test_soup_hits_bit_exact_with_device_builder(cuda_device, n, extent):
    attrs, idx = soup(n, seed=n, extent=extent)
    xf = [IDENTITY, [0.5, 0, 0, 1.2, 0, 0.5, 0, 0.1, 0, 0, 0.5, 0.3]]
    with core.Context(0) as ctx:
        top, _ = build_scene(ctx, attrs, idx, xf, core.BUILD_GPU_LBVH)
        top_host, _ = build_scene(ctx, attrs, idx, xf, core.BUILD_HOST_SAH)
        info = ctx.scene_info(top)
        assert info.numTris == n and info.numNodes >= 2
        rays = H.random_rays(100000, seed=n + 1, lo=(-0.3, -0.3, -0.3), hi=(1.9, 1.3, 1.3), tmin=1e-5)
        got = ctx.trace_closest_host(top, rays)
        assert H.hits_equal(got, ctx.trace_closest_host(top_host, rays))      # builder independence
        ref = oracle_for(attrs, idx, xf)
        assert H.hits_equal(got, ref.trace_closest(rays))
        short = rays.copy()
        short["tmax"] = 0.4
        assert np.array_equal(ctx.trace_any_host(top, short).astype(np.uint8), ref.trace_any(short))
        if n <= 1000:
            assert H.hits_equal(got[:2000], ref.trace_closest(rays[:2000], brute_force=True))
        assert (got["inst"] != 0xffffffff).mean() > (0.001 if n < 10 else 0.2)


def test_degenerate_inputs(cuda_device):
    # all triangles identical (equal Morton codes everywhere), zero-area triangles, and a flat (2D) soup
    with core.Context(0) as ctx:
        attrs, idx = soup(1, seed=5, extent=0.4)
        attrs = np.tile(attrs, 64)
        idx = np.arange(3 * 64, dtype=np.uint32).reshape(64, 3)
        top, _ = build_scene(ctx, attrs, idx, [IDENTITY], core.BUILD_GPU_LBVH)
        ref = oracle_for(attrs, idx, [IDENTITY])
        rays = H.random_rays(20000, seed=9, lo=(-0.2, -0.2, -0.2), hi=(1.2, 1.2, 1.2), tmin=1e-5)
        got = ctx.trace_closest_host(top, rays)
        assert H.hits_equal(got, ref.trace_closest(rays))
        assert (got["prim"][got["inst"] == 0] == 0).all()          # 64 coincident triangles: the lowest id wins every tie

        attrs, idx = soup(5000, seed=6, extent=0.03)
        attrs["vertex"][:, 2] = 0.25                                 # flat soup: zero extent in z
        attrs["vertex"][0:3] = attrs["vertex"][0]                    # one zero-area triangle
        top, _ = build_scene(ctx, attrs, idx, [IDENTITY], core.BUILD_GPU_LBVH)
        ref = oracle_for(attrs, idx, [IDENTITY])
        got = ctx.trace_closest_host(top, rays)
        assert H.hits_equal(got, ref.trace_closest(rays))


def test_bad_index_is_reported(cuda_device):
    attrs, idx = soup(100, seed=1)
    idx[50, 1] = 10 ** 6
    with core.Context(0) as ctx:
        d_a, d_i = ctx.to_device(attrs), ctx.to_device(idx)
        for flags in (core.BUILD_GPU_LBVH, core.BUILD_HOST_SAH):
            with pytest.raises(core.RtcError, match="index out of range"):
                ctx.gas_build(d_a, 48, len(attrs), d_i, len(idx), flags)


def test_rendered_frame_identical_with_device_builder(cuda_device, tmp_path, monkeypatch):
    monkeypatch.setenv("RTC_FORCE_BUILDER", "gpu")
    app = host.App(H.write_system(tmp_path, "rtigo3_geometry", resolution="160 90", samplesSqrt=2), H.scene_path("rtigo3_geometry"))
    try:
        ref = H.oracle_scene(app)
        app.render(4)
        got = app.frame()
        want = ref.render(H.oracle_sys(app), app.info.miss, 160, 90, iter_count=4).reshape(90, 160, 4)
        assert got.tobytes() == want.tobytes()
    finally:
        app.close()
