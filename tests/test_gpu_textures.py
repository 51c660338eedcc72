"""GPU parity, part 6: material textures -- albedo modulation (closesthit.cu:233-240) and the cutout any-hit programs
(anyhit.cu:46-80, :94-132) processed in the canonical candidate order, against the scalar oracle (which is pinned to the
reference's own host-compiled any-hit programs by tests/test_cpu_oracle_vs_reference.py).  Bit-exact frames."""
import numpy as np
import pytest

import helpers as H
from tweeker_raytracer_b200 import core, host

pytestmark = pytest.mark.gpu


def _app(tmp_path, **kw):
    opts = dict(resolution="192 108", samplesSqrt=3)
    opts.update(kw)
    return host.App(H.write_system(tmp_path, "rtigo3_textures", **opts), H.scene_path("rtigo3_textures"))


def _oracle_frame(app, iterations):
    w, h = app.resolution
    ref = H.oracle_scene(app)
    return ref.render(H.oracle_sys(app), app.info.miss, w, h, iter_count=iterations).reshape(h, w, 4)


def test_textured_scene_bit_exact(cuda_device, tmp_path):
    with _app(tmp_path) as app:
        m = app.materials()
        assert (m["textureAlbedo"] != 0).sum() == 3 and (m["textureCutout"] != 0).sum() == 4
        assert app.render(9) == 9
        got = app.frame()
        want = _oracle_frame(app, 9)
        assert got.tobytes() == want.tobytes()
        st = app.stats()
        assert st.stackOverflows == 0 and st.pathSamples == 192 * 108 * 9


@pytest.mark.parametrize("miss,light", [(0, 1), (2, 2)])
def test_textured_scene_other_lights(cuda_device, tmp_path, miss, light):
    """area light only (every shadow ray has finite length) and the HDR environment (importance-sampled env light)."""
    with _app(tmp_path, miss=miss, light=light, envMap="procedural 256 128") as app:
        app.render(4)
        assert app.frame().tobytes() == _oracle_frame(app, 4).tobytes()


def test_deep_paths_and_batches(cuda_device, tmp_path):
    """Russian roulette from the first bounce (every roulette of a path that casts a shadow ray is postponed), long paths,
    iterations split over several enqueues."""
    with _app(tmp_path, pathLengths="0 12", samplesSqrt=3) as app:
        app.render(2); app.render(3); app.render(4)
        assert app.frame().tobytes() == _oracle_frame(app, 9).tobytes()


def test_toggle_textures_like_the_gui(cuda_device, tmp_path):
    with _app(tmp_path, samplesSqrt=2) as app:
        n = app.info.numMaterials
        for i in range(n):
            app.update_material_textures(i, False, False)     # all check boxes off: the plain kernels run again
        assert not app.materials()["textureCutout"].any()
        app.render(4)
        plain = app.frame()
        assert plain.tobytes() == _oracle_frame(app, 4).tobytes()
        app.update_material_textures(1, True, True)           # "floor" was material 1 (0 is the area light's)
        app.update_material_textures(3, False, True)
        app.render(4)
        textured = app.frame()
        assert textured.tobytes() == _oracle_frame(app, 4).tobytes()
        assert textured.tobytes() != plain.tobytes()


def test_opaque_cutout_equals_no_cutout(cuda_device, tmp_path):
    """A cutout texture that is 1 everywhere never draws a random number: the frame equals the one without cutout."""
    import os
    ones = np.full((4, 4, 3), 255, dtype=np.uint8)
    path = os.path.join(str(tmp_path), "ones.ppm")
    with open(path, "wb") as f:
        f.write(b"P6\n4 4\n255\n" + ones.tobytes())
    with _app(tmp_path, samplesSqrt=2, textureCutout=path) as app:
        app.render(4)
        with_cutout = app.frame()
        assert with_cutout.tobytes() == _oracle_frame(app, 4).tobytes()
        for i in range(app.info.numMaterials):
            use_albedo = bool(app.materials()["textureAlbedo"][i] != 0)
            app.update_material_textures(i, use_albedo, False)
        app.render(4)
        assert app.frame().tobytes() == with_cutout.tobytes()
