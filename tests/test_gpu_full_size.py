"""GPU parity at BASELINE.json's full sizes, through properties that do not need the oracle to render a whole frame:
rows of the full-resolution frame (the oracle renders a row subset in seconds), builder independence and idempotence on
a million-triangle soup, closest-hit / any-hit consistency, and conservation of the ray statistics."""
import os
import sys

import numpy as np
import pytest

import helpers as H
from oracle import orc
from tweeker_raytracer_b200 import core, host

pytestmark = pytest.mark.gpu


def test_geometry_scene_1080p_rows_bit_exact(cuda_device, tmp_path):
    # config 2 at its real resolution, 8 of its 256 iterations in one batch; 20 rows (307 200 path samples) against the oracle
    app = host.App(H.write_system(tmp_path, "rtigo3_geometry", samplesSqrt=16), H.scene_path("rtigo3_geometry"))
    try:
        assert app.resolution == (1920, 1080)
        assert app.render(8) == 8
        got = app.frame()
        ref = H.oracle_scene(app)
        st = orc.Stats()
        want = ref.render(H.oracle_sys(app), app.info.miss, 1920, 1080, iter_count=8, row_step=54, row_offset=0, stats=st).reshape(1080, 1920, 4)
        rows = np.arange(0, 1080, 54)
        assert got[rows].tobytes() == want[rows].tobytes()
        assert np.isfinite(got).all() and (got[..., 3] == 1.0).all()
        stats = app.stats()
        assert stats.pathSamples == 1920 * 1080 * 8 and stats.stackOverflows == 0
        # ray counts scale with the row subset: the full frame traces ~54x what the 20 rows do (same scene statistics)
        assert 0.7 < stats.radianceRays / (st.radianceRays * 54.0) < 1.4
    finally:
        app.close()


def test_cornell_config1_full_spec_rows_bit_exact(cuda_device):
    # config 1 exactly as specified: 512x512, 16 spp, fixed seeds; every 16th row against the oracle
    app = host.App(H.SCENES + "/system_rtigo3_cornell_box.txt", H.scene_path("rtigo3_cornell_box"))
    try:
        while app.render(16) < 16:
            pass
        got = app.frame()
        ref = H.oracle_scene(app)
        want = ref.render(H.oracle_sys(app), app.info.miss, 512, 512, iter_count=16, row_step=16).reshape(512, 512, 4)
        assert got[::16].tobytes() == want[::16].tobytes()
    finally:
        app.close()


def test_instanced_stress_scene_config4_full_count_bit_exact(cuda_device, tmp_path):
    # config 4 as specified: 10 000 instances of the 50 000-triangle torus over a floor (500 M instanced triangles, two GAS),
    # at its real resolution; 4 of its 64 iterations; random rays through the lattice and 10 rows of the frame against the oracle
    sys.path.insert(0, os.path.join(H.ROOT, "tools"))
    import make_instances_scene
    scene = os.path.join(str(tmp_path), "scene_rtigo3_instances.txt")
    placed, dims = make_instances_scene.write_scene(scene)
    assert placed == 10000
    app = host.App(H.SCENES + "/system_rtigo3_instances.txt", scene)
    try:
        assert app.resolution == (1920, 1080)
        assert app.info.numInstances == placed + 1 and app.info.numGeometries == 2
        ctx = app.context(0)
        top = app.system_data(0).topObject
        info = ctx.scene_info(top)
        assert info.numGas == 2 and info.numInstances == placed + 1
        assert sum(len(app.geometry(g)[1]) for g in range(2)) >= 50_000          # x 10 000 instances = 500 M instanced triangles
        ref = H.oracle_scene(app)
        half = 0.5 * dims[0] * 2.4
        rays = H.random_rays(200000, seed=4, lo=(-half, 0.05, -half), hi=(half, dims[1] * 2.4 + 2.0, half))
        hits = ctx.trace_closest_host(top, rays)
        assert H.hits_equal(hits, ref.trace_closest(rays))
        assert (hits["inst"] != 0xffffffff).mean() > 0.5
        assert app.render(4) == 4
        got = app.frame()
        want = ref.render(H.oracle_sys(app), app.info.miss, 1920, 1080, iter_count=4, row_step=108, row_offset=31).reshape(1080, 1920, 4)
        rows = np.arange(31, 1080, 108)
        assert got[rows].tobytes() == want[rows].tobytes()
        assert app.stats().stackOverflows == 0
    finally:
        app.close()


def test_million_triangle_soup_properties(cuda_device):
    rng = np.random.default_rng(42)
    n = 1_000_000
    c = rng.uniform(0, 1, size=(n, 1, 3)).astype(np.float32)
    v = (c + rng.uniform(-0.01, 0.01, size=(n, 3, 3)).astype(np.float32)).reshape(-1, 3)
    idx = np.arange(3 * n, dtype=np.uint32)
    inst = np.zeros(1, dtype=core.INSTANCE_DTYPE)
    inst[0]["transform"] = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]
    with core.Context(0) as ctx:
        d_v, d_i = ctx.to_device(v), ctx.to_device(idx)
        tops = []
        for flags in (core.BUILD_GPU_LBVH, core.BUILD_HOST_SAH):
            inst[0]["gas"] = ctx.gas_build(d_v, 12, 3 * n, d_i, n, flags)
            tops.append(ctx.ias_build(inst))
        rays = H.random_rays(1_000_000, seed=99, lo=(0, 0, 0), hi=(1, 1, 1), tmin=1e-5)
        a = ctx.trace_closest_host(tops[0], rays)
        b = ctx.trace_closest_host(tops[1], rays)
        assert H.hits_equal(a, b)                                   # independent of the acceleration structure
        assert H.hits_equal(a, ctx.trace_closest_host(tops[0], rays))   # idempotent
        hit = a["inst"] == 0
        assert 0.9 < hit.mean() <= 1.0 and (a["prim"][hit] < n).all() and (a["t"][hit] > 1e-5).all()
        assert ((a["u"][hit] >= 0) & (a["v"][hit] >= 0) & (a["u"][hit] + a["v"][hit] <= 1.0 + 1e-5)).all()
        # any-hit with tmax just beyond / just before the closest hit
        beyond, before = rays.copy(), rays.copy()
        beyond["tmax"] = np.where(hit, np.nextafter(a["t"], np.float32(np.inf)), rays["tmax"])
        before["tmax"] = np.where(hit, a["t"], rays["tmax"])          # the interval is open: the closest hit itself is excluded
        assert np.array_equal(ctx.trace_any_host(tops[0], beyond) != 0, hit)
        assert not ctx.trace_any_host(tops[0], before)[hit].any()
        # a sample against the brute-force oracle (ground truth for the intersector arithmetic)
        s = orc.Scene()
        attrs = np.zeros(3 * n, dtype=orc.ATTR_DTYPE)
        attrs["vertex"] = v
        g = s.add_geometry(attrs, idx.reshape(-1, 3))
        s.add_instance(inst[0]["transform"], g, 0, -1)
        s.set_materials(np.zeros(1, dtype=orc.MATERIAL_DTYPE))
        s.set_lights(np.zeros(0, dtype=orc.LIGHT_DTYPE))
        s.set_camera(np.zeros(1, dtype=orc.CAMERA_DTYPE))
        s.commit()
        assert H.hits_equal(a[:20000], s.trace_closest(rays[:20000]))
        assert ctx.stats().stackOverflows == 0
