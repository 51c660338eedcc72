"""GPU parity, part 4: the two-level path under many instances (BASELINE config 4, scaled down) and the in-process
multi-GPU strategies of the Raytracer classes (need >= 2 GPUs; skipped otherwise)."""
import os
import sys

import numpy as np
import pytest

import helpers as H
from oracle import orc
from tweeker_raytracer_b200 import core, host, partition

sys.path.insert(0, os.path.join(H.ROOT, "tools"))
import make_instances_scene  # noqa: E402

pytestmark = pytest.mark.gpu


def test_instanced_scene_bit_exact(cuda_device, tmp_path):
    scene = os.path.join(str(tmp_path), "scene_instances.txt")
    placed, _ = make_instances_scene.write_scene(scene, count=400, tess=(48, 24))
    app = host.App(H.write_system(tmp_path, "rtigo3_instances", resolution="256 144", samplesSqrt=2, camera="0.8 0.62 55 32"), scene)
    try:
        assert app.info.numInstances == placed + 1 and app.info.numGeometries == 2      # one shared torus GAS + the floor
        ref = H.oracle_scene(app)
        ctx = app.context(0)
        top = app.system_data(0).topObject
        info = ctx.scene_info(top)
        assert info.numGas == 2 and info.numInstances == placed + 1
        rays = H.random_rays(150000, seed=77, lo=(-25, 0.05, -25), hi=(25, 30, 25))
        assert H.hits_equal(ctx.trace_closest_host(top, rays), ref.trace_closest(rays))
        app.render(4)
        got = app.frame()
        want = ref.render(H.oracle_sys(app), app.info.miss, 256, 144, iter_count=4).reshape(144, 256, 4)
        assert got.tobytes() == want.tobytes()
    finally:
        app.close()


def oracle_tiled(app, count, iterations):
    """Reference result of the tiled strategies: per-device local-copy renders composited (raygeneration.cu:259-344, compositor.cu)."""
    ref = H.oracle_scene(app)
    w, h = app.resolution
    base = H.oracle_sys(app)
    lw = partition.tiled_launch_width(w, count, base.tileSize.x)
    out = np.zeros((h, w, 4), dtype=np.float32)
    for index in range(count):
        sysd = H.oracle_sys(app)
        sysd.deviceCount, sysd.deviceIndex, sysd.distribution = count, index, 1
        slab = ref.render(sysd, app.info.miss, lw, h, local_copy=True, iter_count=iterations).reshape(h, lw, 4)
        args = orc.CompositorData()
        args.resolution.x, args.resolution.y = w, h
        args.tileSize.x, args.tileSize.y, args.tileShift.x, args.tileShift.y = sysd.tileSize.x, sysd.tileSize.y, sysd.tileShift.x, sysd.tileShift.y
        args.launchWidth, args.deviceCount, args.deviceIndex = lw, count, index
        orc.composite(args, slab, out)
    return out


@pytest.mark.parametrize("strategy,composite", [(1, 0), (2, 0), (3, 0), (3, 1)])
def test_multi_gpu_strategies_bit_exact(cuda_device, tmp_path, strategy, composite):
    if cuda_device < 2:
        pytest.skip("needs two GPUs")
    n = min(cuda_device, 4)
    app = host.App(H.write_system(tmp_path, "rtigo3_cornell_box", resolution="120 64", samplesSqrt=2, strategy=strategy,
                                  devicesMask=(1 << n) - 1, composite=composite), H.scene_path("rtigo3_cornell_box"))
    try:
        assert app.info.numDevices == n and app.info.strategy == strategy
        assert app.render(3) == 3
        got = app.frame()
        want = oracle_tiled(app, n, 3)
        assert got.tobytes() == want.tobytes()
        stats = app.stats()
        assert stats.pathSamples == 120 * 64 * 3
    finally:
        app.close()
