"""Known answers that follow from the reference's code, checked on the oracle without a GPU."""
import numpy as np
import pytest

import helpers as H
from oracle import orc
from tweeker_raytracer_b200 import host


def test_white_furnace(built, tmp_path):
    scene = tmp_path / "scene_furnace.txt"
    scene.write_text("albedo 1 1 1\nmaterial default brdf_diffuse\nmaterial white brdf_diffuse\nidentity\nmodel sphere 48 24 1.0 white\n")
    sysfile = H.write_system(tmp_path, "rtigo3_cornell_box", resolution="24 24", samplesSqrt=16, miss=1, light=0,
                             pathLengths="2 2", center="0 0 0", camera="0.75 0.5 45 4")
    app = host.App(sysfile, str(scene), host_only=True)
    img = H.oracle_scene(app).render(H.oracle_sys(app), 1, 24, 24, iter_count=256)[:, :3]
    assert abs(float(img.mean()) - 1.0) < 0.01
    app.close()


def test_no_lights_is_black_and_traces_no_shadow_rays(built, tmp_path):
    app = host.App(H.write_system(tmp_path, "rtigo3_cornell_box", resolution="16 16", light=0), H.scene_path("rtigo3_cornell_box"), host_only=True)
    st = orc.Stats()
    img = H.oracle_scene(app).render(H.oracle_sys(app), 0, 16, 16, iter_count=2, stats=st)
    assert float(img[:, :3].max()) == 0.0 and st.shadowRays == 0 and st.pathSamples == 512 and (img[:, 3] == 1).all()
    app.close()


def test_running_average_matches_iterating_one_by_one(built, tmp_path):
    app = host.App(H.write_system(tmp_path, "rtigo3_cornell_box", resolution="12 12"), H.scene_path("rtigo3_cornell_box"), host_only=True)
    s, sysd = H.oracle_scene(app), H.oracle_sys(app)
    a = s.render(sysd, 0, 12, 12, iter_count=5)
    b = np.zeros((144, 4), dtype=np.float32)
    for it in range(5):
        s.render(sysd, 0, 12, 12, iter_first=it, iter_count=1, buffer=b)
    assert a.tobytes() == b.tobytes()
    # row subsets (the bounded CPU baseline sample) only touch their rows
    c = s.render(sysd, 0, 12, 12, iter_count=5, row_step=4, row_offset=1).reshape(12, 12, 4)
    full = a.reshape(12, 12, 4)
    assert np.array_equal(c[1::4], full[1::4]) and float(np.abs(c[0::4]).max()) == 0.0
    app.close()


def test_power_heuristic_and_tonemap_known_values(built):
    # tonemapper with neutral parameters is the gamma curve only (Application.cpp:2262-2295)
    p = orc.TonemapperParams(1.0, (1, 1, 1), 1.0, 1.0, 0.0, 1.0, 1.0)
    rgba = np.array([[0, 0, 0, 1], [0.5, 0.5, 0.5, 1], [1, 1, 1, 1], [4, 0.25, 0, 1]], dtype=np.float32)
    out = orc.tonemap(p, rgba)
    assert out[0].tolist() == [0, 0, 0] and out[2].tolist() == [255, 255, 255] and out[1].tolist() == [127, 127, 127]
    assert out[3, 0] == 255 and out[3, 2] == 0
    p2 = orc.TonemapperParams(2.2, (1, 1, 1), 1.0, 0.8, 0.2, 1.2, 0.8)
    o2 = orc.tonemap(p2, rgba)
    assert o2[1, 0] > o2[0, 0] and (o2[2] >= o2[1]).all()


def test_compositor_scatter(built):
    from tweeker_raytracer_b200 import partition
    w, h, n, tile = 20, 6, 3, 4
    lw = partition.tiled_launch_width(w, n, tile)
    out = np.full((h, w, 4), -1, dtype=np.float32)
    for d in range(n):
        slab = np.zeros((h, lw, 4), dtype=np.float32)
        for y in range(h):
            for x in range(lw):
                slab[y, x] = (d, x, y, 1)
        args = orc.CompositorData()
        args.resolution.x, args.resolution.y, args.tileSize.x, args.tileSize.y = w, h, tile, tile
        args.tileShift.x = args.tileShift.y = 2
        args.launchWidth, args.deviceCount, args.deviceIndex = lw, n, d
        orc.composite(args, slab, out)
    assert (out[..., 3] == 1).all()
    for y in range(h):
        for col in range(w):
            d, x = int(out[y, col, 0]), int(out[y, col, 1])
            assert partition.distribute(x, y, d, n, tile, 2, 2) == col and out[y, col, 2] == y
