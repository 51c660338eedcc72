#!/usr/bin/env python
"""Generates tests/golden/*.npz / *.json FROM THE REFERENCE ITSELF: the reference's shader sources are compiled for the
host where they lie under /root/reference (oracle/Makefile target `ref`, nothing is copied) and run by
oracle/ref_driver.cpp.  Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H  # noqa: E402
from oracle import orc  # noqa: E402
from tweeker_raytracer_b200 import host  # noqa: E402

CASES = {
    "cornell_32x32_4spp": ("rtigo3_cornell_box", dict(resolution="32 32", samplesSqrt=2), 4),
    "cornell_fisheye_24x24_2spp": ("rtigo3_cornell_box", dict(resolution="24 24", samplesSqrt=2, lensShader=1), 2),
    "cornell_sphere_24x24_2spp": ("rtigo3_cornell_box", dict(resolution="24 24", samplesSqrt=2, lensShader=2), 2),
    "geometry_48x27_4spp": ("rtigo3_geometry", dict(resolution="48 27", samplesSqrt=2), 4),
    "geometry_env_48x27_4spp": ("rtigo3_geometry", dict(resolution="48 27", samplesSqrt=2, miss=2, envMap="procedural 128 64", envRotation=0.15), 4),
    # material textures: albedo modulation + the reference's own cutout any-hit programs run per candidate in canonical order
    "textures_64x36_4spp": ("rtigo3_textures", dict(resolution="64 36", samplesSqrt=2), 4),
    "textures_rr_env_48x27_3spp": ("rtigo3_textures", dict(resolution="48 27", samplesSqrt=2, miss=2, envMap="procedural 128 64", pathLengths="0 8"), 3),
    # converged frames for the PSNR >= 40 dB check of BASELINE.json (the host-compiled reference stands in for the OptiX build)
    "geometry_converged_96x54_256spp": ("rtigo3_geometry", dict(resolution="96 54", samplesSqrt=16), 256),
    "cornell_converged_64x64_256spp": ("rtigo3_cornell_box", dict(resolution="64 64", samplesSqrt=16), 256),
    "cornell_tiled_3dev_40x16_2spp": ("rtigo3_cornell_box", dict(resolution="40 16", samplesSqrt=2, tileSize="8 8"), 2),
}


def main():
    tmp = tempfile.mkdtemp()
    frames = {}
    ref0 = None
    for key, (name, overrides, iterations) in CASES.items():
        app = host.App(H.write_system(tmp, name, **overrides), H.scene_path(name), host_only=True)
        scene = H.oracle_scene(app, "libm")
        ref = orc.Reference(scene, app.info.miss)
        ref0 = ref0 or ref
        sysd = H.oracle_sys(app)
        w, h = app.resolution
        if "tiled" in key:
            from tweeker_raytracer_b200 import partition
            lw = partition.tiled_launch_width(w, 3, sysd.tileSize.x)
            for index in range(3):
                sysd.deviceCount, sysd.deviceIndex, sysd.distribution = 3, index, 1
                frames["%s_dev%d" % (key, index)] = ref.render(sysd, lw, h, local_copy=True, iter_count=iterations).reshape(h, lw, 4)
        else:
            frames[key] = ref.render(sysd, w, h, iter_count=iterations).reshape(h, w, 4)
        app.close()
    np.savez_compressed(os.path.join(HERE, "reference_frames.npz"), **frames)
    rng = {"tea4": [[a, b, ref0.tea4(a, b)] for a, b in [(0, 0), (1, 0), (0, 1), (12345, 7), (0xffffffff, 0xffffffff), (512 * 100 + 37, 15)]],
           "lcg": []}
    for seed in (0, 1, 0x9e3779b9, 0xdeadbeef):
        seq, state = ref0.rng_sequence(seed, 8)
        rng["lcg"].append({"seed": seed, "samples_hex": [float(np.float32(x)).hex() for x in seq], "state": state})
    with open(os.path.join(HERE, "reference_rng.json"), "w") as f:
        json.dump(rng, f, indent=1)
    print("wrote", sorted(frames), "and reference_rng.json")


if __name__ == "__main__":
    main()
