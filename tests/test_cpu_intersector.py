"""The ray/triangle arithmetic this repository DEFINES (the reference keeps it in OptiX): properties it must have."""
import numpy as np
import pytest

import helpers as H
from oracle import orc
from tweeker_raytracer_b200 import host

IDENTITY = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]


def grid_scene(n=8, transform=IDENTITY):
    """(n x n) cells in the plane z = 0 over [-1,1]^2, two triangles per cell."""
    xs = np.linspace(-1, 1, n + 1, dtype=np.float32)
    attrs = np.zeros((n + 1) * (n + 1), dtype=orc.ATTR_DTYPE)
    for j in range(n + 1):
        for i in range(n + 1):
            attrs[j * (n + 1) + i]["vertex"] = (xs[i], xs[j], 0)
            attrs[j * (n + 1) + i]["normal"] = (0, 0, 1)
            attrs[j * (n + 1) + i]["tangent"] = (1, 0, 0)
    idx = []
    for j in range(n):
        for i in range(n):
            a = j * (n + 1) + i
            idx += [a, a + 1, a + n + 2, a + n + 2, a + n + 1, a]
    s = orc.Scene()
    g = s.add_geometry(attrs, np.array(idx, dtype=np.uint32))
    s.add_instance(transform, g, 0, -1)
    s.set_materials(np.zeros(1, dtype=orc.MATERIAL_DTYPE))
    s.set_lights(np.zeros(0, dtype=orc.LIGHT_DTYPE))
    s.set_camera(np.zeros(1, dtype=orc.CAMERA_DTYPE))
    s.commit()
    return s, xs


def rays_to(points, origin=(0.1, -0.2, 3.0)):
    rays = np.zeros(len(points), dtype=orc.RAY_DTYPE)
    o = np.array(origin, dtype=np.float32)
    d = (np.asarray(points, dtype=np.float32) - o)
    rays["ox"], rays["oy"], rays["oz"] = o
    rays["dx"], rays["dy"], rays["dz"] = d[:, 0], d[:, 1], d[:, 2]
    rays["tmin"], rays["tmax"] = 1e-4, 1e27
    return rays


def test_watertight_on_shared_edges_and_vertices(built):
    s, xs = grid_scene(8)
    pts = []
    for x in xs[1:-1]:                       # interior grid lines: shared edges
        for t in np.linspace(-0.95, 0.95, 41):
            pts.append((x, t, 0))
            pts.append((t, x, 0))
    for x in xs[1:-1]:                       # interior vertices: up to six triangles meet
        for y in xs[1:-1]:
            pts.append((x, y, 0))
    for t in np.linspace(-0.99, 0.99, 97):   # cell diagonals
        pts.append((t, t, 0))
    rays = rays_to(pts)
    hits = s.trace_closest(rays)
    assert (hits["inst"] == 0).all(), "a ray through a shared edge or vertex slipped between two triangles"
    assert np.array_equal(hits.tobytes(), s.trace_closest(rays, brute_force=True).tobytes())
    assert np.allclose(hits["t"], 1.0, atol=1e-5)      # the target points lie at parameter 1 of un-normalised directions


def test_bvh_equals_brute_force_on_random_rays(built, tmp_path):
    app = host.App(H.write_system(tmp_path, "rtigo3_cornell_box", resolution="8 8"), H.scene_path("rtigo3_cornell_box"), host_only=True)
    s = H.oracle_scene(app)
    rays = H.random_rays(1500, seed=11, lo=(-1, 0, -1), hi=(1, 2, 1))
    assert s.trace_closest(rays).tobytes() == s.trace_closest(rays, brute_force=True).tobytes()
    short = H.random_rays(1500, seed=12, lo=(-1, 0, -1), hi=(1, 2, 1), tmax=0.7)
    assert np.array_equal(s.trace_any(short), s.trace_any(short, brute_force=True))
    app.close()


def test_interval_is_open_and_miss_is_encoded(built):
    s, _ = grid_scene(2)
    r = np.zeros(4, dtype=orc.RAY_DTYPE)
    r[0] = (0.3, 0.3, 2.0, 0.0, 0, 0, -1, 1e27)      # hits at t = 2
    r[1] = (0.3, 0.3, 2.0, 2.0, 0, 0, -1, 1e27)      # tmin == t: excluded
    r[2] = (0.3, 0.3, 2.0, 0.0, 0, 0, -1, 2.0)       # tmax == t: excluded
    r[3] = (0.3, 0.3, 2.0, 0.0, 0, 0, 1, 1e27)       # points away
    h = s.trace_closest(r)
    assert h["t"][0] == 2.0 and h["inst"][0] == 0
    for k in (1, 2, 3):
        assert h["inst"][k] == 0xffffffff and h["prim"][k] == 0xffffffff and h["t"][k] == -1.0
    assert list(s.trace_any(r)) == [1, 0, 0, 0]


def test_both_faces_hit_and_barycentrics_weight_v1_v2(built):
    s, _ = grid_scene(1)
    r = np.zeros(2, dtype=orc.RAY_DTYPE)
    r[0] = (0.5, -0.5, 1.0, 0.0, 0, 0, -1, 1e27)     # front
    r[1] = (0.5, -0.5, -1.0, 0.0, 0, 0, 1, 1e27)     # back: no culling (Device.cpp:1378)
    h = s.trace_closest(r)
    assert (h["inst"] == 0).all() and h["prim"][0] == h["prim"][1] == 0
    # triangle 0 = (-1,-1) (1,-1) (1,1): point (0.5,-0.5) = v0 + 0.75 (v1-v0) + 0.25 (v2-v1) -> beta 0.5, gamma 0.25
    assert np.allclose(h["u"], 0.5) and np.allclose(h["v"], 0.25)


def test_equal_t_ties_go_to_the_lower_instance_then_primitive(built):
    # two coincident instances of the same quad: every hit has an exact tie
    s, _ = grid_scene(2)
    s2 = orc.Scene()
    attrs, idx = s.keep["geometries"][0]
    g = s2.add_geometry(attrs, idx)
    s2.add_instance(IDENTITY, g, 0, -1)
    s2.add_instance(IDENTITY, g, 0, -1)
    s2.set_materials(np.zeros(1, dtype=orc.MATERIAL_DTYPE))
    s2.set_lights(np.zeros(0, dtype=orc.LIGHT_DTYPE))
    s2.set_camera(np.zeros(1, dtype=orc.CAMERA_DTYPE))
    s2.commit()
    rays = H.random_rays(500, seed=3, lo=(-0.9, -0.9, 0.5), hi=(0.9, 0.9, 2.0))
    rays["dx"], rays["dy"], rays["dz"] = 0, 0, -1
    h = s2.trace_closest(rays)
    assert (h["inst"] == 0).all()
    assert h.tobytes() == s2.trace_closest(rays, brute_force=True).tobytes()


def test_instance_transform_shares_t_with_the_world_ray(built):
    # scale 0.25 + translate: the object-space ray is not normalised, so t is the world-space parameter
    xf = [0.25, 0, 0, 2.0, 0, 0.25, 0, -1.0, 0, 0, 0.25, 0.5]
    s, _ = grid_scene(4, xf)
    r = np.zeros(1, dtype=orc.RAY_DTYPE)
    r[0] = (2.05, -1.02, 3.5, 0.0, 0, 0, -1, 1e27)
    h = s.trace_closest(r)
    assert h["inst"][0] == 0 and abs(h["t"][0] - 3.0) < 1e-6
    inv = s.inverse(0)
    assert np.allclose(inv, [4, 0, 0, -8, 0, 4, 0, 4, 0, 0, 4, -2])
