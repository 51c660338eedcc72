"""GPU parity, part 7: an imported Wavefront OBJ mesh (`model assimp`), instanced as a whole model on a ring like the
reference's scene_rtigo3_instances.txt: closest hits and rendered frame bit-exact against the oracle."""
import os
import sys

import pytest

import helpers as H
from tweeker_raytracer_b200 import host

sys.path.insert(0, os.path.join(H.ROOT, "tools"))
import make_obj_mesh  # noqa: E402

pytestmark = pytest.mark.gpu


def test_mesh_ring_bit_exact(cuda_device, tmp_path):
    scene, triangles = make_obj_mesh.write_scene(str(tmp_path), copies=9, subdivisions=4)
    assert triangles == 20 * 4 ** 4
    system = H.write_system(tmp_path, "rtigo3_geometry", resolution="224 126", samplesSqrt=2, camera="0.6 0.5 55 14", center="0 1 0")
    with host.App(system, scene) as app:
        assert app.info.numGeometries == 1 + 1 + 2           # light quad, floor, blob top + bottom (shared by all copies)
        assert app.info.numInstances == 1 + 1 + 2 * 9
        ref = H.oracle_scene(app)
        ctx = app.context(0)
        top = app.system_data(0).topObject
        assert ctx.scene_info(top).numGas == 4
        rays = H.random_rays(120000, seed=5, lo=(-7, 0.05, -7), hi=(7, 4, 7))
        assert H.hits_equal(ctx.trace_closest_host(top, rays), ref.trace_closest(rays))
        app.render(4)
        want = ref.render(H.oracle_sys(app), app.info.miss, 224, 126, iter_count=4).reshape(126, 224, 4)
        assert app.frame().tobytes() == want.tobytes()
