#!/usr/bin/env python
"""Generates tests/golden/reference_envcdf.npz FROM THE REFERENCE ITSELF: apps/rtigo3/src/Texture.cpp is compiled for the host
where it lies (oracle/Makefile target `reftex`, nothing is copied) and its calculateSphericalCDF (Texture.cpp:1540-1645) is
run on the texels below.  Run in the build container (needs /root/reference):  python tests/golden/make_golden_envcdf.py

Cases: the host's procedural sky at two sizes (the texels come from host/EnvMap.cpp and are stored, so the fixture also
pins them), and a seeded random Radiance .hdr file (helpers.random_rgbe: a black row, a black column band, large dynamic range) read by
the host's loader."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H  # noqa: E402
from tweeker_raytracer_b200 import host  # noqa: E402


def reference_cdf(rgba):
    """(cdf_u [h, w+1], cdf_v [h+1], integral) by the reference's own Texture::calculateSphericalCDF."""
    lib = os.path.join(ROOT, "oracle", "_ref", "libreftex.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "reftex"])
    L = C.CDLL(lib)
    t = np.ascontiguousarray(rgba, dtype=np.float32)
    h, w = t.shape[:2]
    cdf_u = np.zeros((h, w + 1), dtype=np.float32)
    cdf_v = np.zeros(h + 1, dtype=np.float32)
    integral = C.c_float(0)
    L.reftex_spherical_cdf(t.ctypes.data_as(C.c_void_p), w, h, cdf_u.ctypes.data_as(C.c_void_p), cdf_v.ctypes.data_as(C.c_void_p), C.byref(integral))
    return cdf_u, cdf_v, np.float32(integral.value)


def main():
    tmp = tempfile.mkdtemp()
    out = {}
    for key, spec in (("procedural_64x32", "procedural 64 32"), ("procedural_256x128", "procedural 256 128")):
        app = host.App(H.write_system(tmp, "rtigo3_geometry", miss=2, envMap=spec, resolution="8 8"), H.scene_path("rtigo3_geometry"), host_only=True)
        texels = app.environment()[0]
        app.close()
        u, v, i = reference_cdf(texels)
        out[key + "_texels"], out[key + "_cdf_u"], out[key + "_cdf_v"], out[key + "_integral"] = texels, u, v, i
    # a Radiance .hdr FILE through the host's reader (Texture.cpp:1300-1377 reads it with DevIL): texels, then the reference's CDF
    path = H.write_rgbe_hdr(os.path.join(tmp, "random.hdr"), H.random_rgbe(48, 24, 20261018))
    app = host.App(H.write_system(tmp, "rtigo3_geometry", miss=2, envMap=path, resolution="8 8"), H.scene_path("rtigo3_geometry"), host_only=True)
    texels = app.environment()[0]
    app.close()
    u, v, i = reference_cdf(texels)
    out["random_48x24_texels"], out["random_48x24_cdf_u"], out["random_48x24_cdf_v"], out["random_48x24_integral"] = texels, u, v, i
    np.savez_compressed(os.path.join(HERE, "reference_envcdf.npz"), **out)
    print("wrote reference_envcdf.npz:", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
