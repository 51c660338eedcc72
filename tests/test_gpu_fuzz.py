"""GPU parity, part 9: randomly generated scenes (fixed seeds) -- random instance transforms with non-uniform scale and
rotation, all model kinds, all BSDFs with random parameters incl. nested transmissive volumes, random light / miss / lens
set-ups and cameras inside or outside the geometry.  Closest hits of random rays and the rendered frame must equal the
oracle bit for bit."""
import os

import numpy as np
import pytest

import helpers as H
from tweeker_raytracer_b200 import host

pytestmark = pytest.mark.gpu

BSDFS = ["brdf_diffuse", "brdf_specular", "bsdf_specular", "brdf_ggx_smith", "bsdf_ggx_smith"]


def random_scene(path, rng, textures=False):
    lines = ["albedo 0.5 0.5 0.5", "material default brdf_diffuse"]
    names = ["default"]
    for m in range(int(rng.integers(3, 8))):
        a = rng.uniform(0.05, 1.0, 3)
        r = rng.uniform(0.02, 0.8, 2)
        c = rng.uniform(0.1, 1.0, 3)
        lines += ["albedo %.4f %.4f %.4f" % tuple(a), "roughness %.4f %.4f" % tuple(r), "absorption %.4f %.4f %.4f" % tuple(c),
                  "absorptionScale %.3f" % (rng.uniform(0.0, 3.0) if rng.random() < 0.5 else 0.0), "ior %.3f" % rng.uniform(1.05, 2.2),
                  "thinwalled %d" % int(rng.random() < 0.25)]
        if textures:
            lines += ["albedoTexture %d" % int(rng.random() < 0.4), "cutoutTexture %d" % int(rng.random() < 0.5)]
        lines += ["material m%d %s" % (m, BSDFS[int(rng.integers(0, 5))])]
        names.append("m%d" % m)
    lines.append("identity")
    lines.append("push scale 8 1 8 model plane %d %d 1 default pop" % (int(rng.integers(1, 5)), int(rng.integers(1, 5))))
    for _ in range(int(rng.integers(4, 14))):
        kind = int(rng.integers(0, 4))
        model = ["box", "sphere %d %d %.2f" % (int(rng.integers(6, 40)), int(rng.integers(4, 20)), rng.choice([1.0, 0.5, 0.75])),
                 "torus %d %d %.2f %.2f" % (int(rng.integers(6, 40)), int(rng.integers(6, 30)), rng.uniform(0.5, 1.0), rng.uniform(0.1, 0.4)),
                 "plane %d %d %d" % (int(rng.integers(1, 4)), int(rng.integers(1, 4)), int(rng.integers(0, 3)))][kind]
        axis = rng.normal(size=3)
        s = rng.uniform(0.2, 1.6, 3)
        t = rng.uniform([-4, 0.2, -4], [4, 3.0, 4])
        lines.append("push scale %.3f %.3f %.3f rotate %.3f %.3f %.3f %.1f translate %.3f %.3f %.3f model %s %s pop"
                     % (s[0], s[1], s[2], axis[0], axis[1], axis[2], rng.uniform(0, 360), t[0], t[1], t[2], model, names[int(rng.integers(0, len(names)))]))
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


@pytest.mark.parametrize("seed", list(range(16)))
def test_random_scene_bit_exact(cuda_device, tmp_path, seed):
    rng = np.random.default_rng(1000 + seed)
    scene = os.path.join(str(tmp_path), "scene_fuzz.txt")
    random_scene(scene, rng)
    miss = int(rng.integers(0, 3))
    overrides = dict(resolution="%d %d" % (int(rng.integers(20, 90)), int(rng.integers(12, 60))), samplesSqrt=2, miss=miss,
                     light=int(rng.integers(0 if miss else 1, 3)), lensShader=int(rng.integers(0, 3)),
                     pathLengths="%d %d" % (int(rng.integers(0, 4)), int(rng.integers(1, 12))),
                     camera="%.3f %.3f %.1f %.2f" % (rng.uniform(0, 1), rng.uniform(0.3, 0.7), rng.uniform(30, 90), rng.uniform(1.5, 14)),
                     center="%.2f %.2f %.2f" % tuple(rng.uniform([-1, 0.5, -1], [1, 2, 1])), envMap="procedural 64 32",
                     envRotation="%.3f" % rng.uniform(0, 1), epsilonFactor=int(rng.choice([100, 500, 2000])))
    with host.App(H.write_system(tmp_path, "rtigo3_geometry", **overrides), scene) as app:
        w, h = app.resolution
        ref = H.oracle_scene(app)
        ctx = app.context(0)
        top = app.system_data(0).topObject
        rays = H.random_rays(30000, seed=seed, lo=(-5, 0.05, -5), hi=(5, 4, 5))
        assert H.hits_equal(ctx.trace_closest_host(top, rays), ref.trace_closest(rays))
        app.render(4)
        want = ref.render(H.oracle_sys(app), app.info.miss, w, h, iter_count=4).reshape(h, w, 4)
        assert app.frame().tobytes() == want.tobytes()
        assert app.stats().stackOverflows == 0


@pytest.mark.parametrize("seed", list(range(12)))
def test_random_textured_scene_bit_exact(cuda_device, tmp_path, seed):
    """The same with albedo and cutout textures switched on at random: ordered any-hit rounds on radiance and shadow rays,
    postponed Russian roulette, overlapping cutout surfaces."""
    rng = np.random.default_rng(5000 + seed)
    scene = os.path.join(str(tmp_path), "scene_fuzz_tex.txt")
    random_scene(scene, rng, textures=True)
    miss = int(rng.integers(0, 3))
    overrides = dict(resolution="%d %d" % (int(rng.integers(24, 80)), int(rng.integers(16, 48))), samplesSqrt=2, miss=miss,
                     light=int(rng.integers(0 if miss else 1, 3)), pathLengths="%d %d" % (int(rng.integers(0, 3)), int(rng.integers(2, 10))),
                     camera="%.3f %.3f %.1f %.2f" % (rng.uniform(0, 1), rng.uniform(0.3, 0.7), rng.uniform(35, 80), rng.uniform(3, 12)),
                     center="0 1 0", envMap="procedural 64 32")
    with host.App(H.write_system(tmp_path, "rtigo3_textures", **overrides), scene) as app:
        w, h = app.resolution
        ref = H.oracle_scene(app)
        app.render(4)
        want = ref.render(H.oracle_sys(app), app.info.miss, w, h, iter_count=4).reshape(h, w, 4)
        assert app.frame().tobytes() == want.tobytes()
