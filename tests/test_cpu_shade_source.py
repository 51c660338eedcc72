"""The PRODUCT's ray-generation source on the CPU: csrc/shade.cuh compiled unchanged by g++ (tests/native/shade_host.cpp, the
flags of the device build of the shading translation unit: no FMA contraction, IEEE division and square root) and held against

  * the reference: its TEA and LCG generators reproduce the golden vectors made by the reference's own sources
    (tests/golden/reference_rng.json);
  * the oracle (which is pinned against the reference bit for bit): start_path() -- distribute(), seeding, jitter, the pinhole /
    fisheye / sphere lens shaders -- gives the oracle's primary rays bit for bit, for a single device and for every device of a
    tiled multi-GPU launch, where launch indices outside the image must be skipped identically;
  * the texture fetch this repository defines (bilinear, wrap / wrap): equal to the oracle's, bit for bit, inside and far
    outside the unit square.

  * the BSDF, light and miss callables: the five sample callables, the eval callables, the constant / spherical environment and
    parallelogram lights and the three miss programs equal the oracle's restatement word for word (function-level test hooks
    of the oracle, orc_test_*).

This is the CPU twin of tests/test_gpu_trace_parity.py's primary-ray test and of what the frame tests prove about the callables;
the closest-hit glue and the integrator are only reachable through whole frames and stay GPU tests (tests/test_gpu_render_parity.py)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from oracle import orc
from tweeker_raytracer_b200 import host

GOLDEN = os.path.join(H.ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def shade():
    src = os.path.join(H.ROOT, "tests", "native", "shade_host.cpp")
    out_dir = os.path.join(H.ROOT, "oracle", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libshade_host.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I" + os.path.join(H.ROOT, "include"),
                           "-I" + os.path.join(H.ROOT, "tweeker_raytracer_b200", "csrc"), "-I/usr/local/cuda/include", "-o", so, src])
    L = C.CDLL(so)
    L.sh_tea4.argtypes = [C.c_uint32, C.c_uint32]
    L.sh_tea4.restype = C.c_uint32
    L.sh_rng_sequence.argtypes = [C.c_uint32, C.c_int, C.c_void_p, C.POINTER(C.c_uint32)]
    L.sh_generate_primary.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
    L.sh_tex2d.argtypes = [C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
    return L


def test_generators_reproduce_the_reference_golden_vectors(built, shade):
    with open(os.path.join(GOLDEN, "reference_rng.json")) as f:
        g = json.load(f)
    for a, b, want in g["tea4"]:
        assert shade.sh_tea4(a, b) == want
    for row in g["lcg"]:
        n = len(row["samples_hex"])
        seq = np.zeros(n, dtype=np.float32)
        state = C.c_uint32(0)
        shade.sh_rng_sequence(row["seed"], n, seq.ctypes.data, C.byref(state))
        assert state.value == row["state"]
        assert [float(x).hex() for x in seq] == row["samples_hex"]
    rng = np.random.default_rng(3)
    for a, b in rng.integers(0, 2 ** 32, size=(200, 2), dtype=np.uint64):
        assert shade.sh_tea4(int(a), int(b)) == orc.tea4(int(a), int(b))


def primary_rays(shade, app, device_index, launch_width, launch_height, iteration):
    sysd = app.system_data(device_index)
    camera = np.ascontiguousarray(app.camera())
    sysd.cameraDefinitions = camera.ctypes.data          # host_only: the host copy stands in for the device array
    rays = np.zeros(launch_width * launch_height, dtype=orc.RAY_DTYPE)
    shade.sh_generate_primary(C.byref(sysd), launch_width, launch_height, iteration, rays.ctypes.data)
    return rays


def rays_identical(a, b):
    return all(np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)) for k in a.dtype.names)


@pytest.mark.parametrize("lens", [0, 1, 2])
def test_primary_rays_equal_the_oracle_for_every_lens_shader(built, shade, tmp_path, lens):
    app = host.App(H.write_system(tmp_path, "rtigo3_cornell_box", resolution="96 64", samplesSqrt=2, lensShader=lens),
                   H.scene_path("rtigo3_cornell_box"), host_only=True)
    ref = H.oracle_scene(app)
    for iteration in (0, 3):
        got = primary_rays(shade, app, 0, 96, 64, iteration)
        want = ref.generate_primary(H.oracle_sys(app), 96, 64, iteration)
        assert rays_identical(got, want)
        assert (got["tmax"] > 0).all()
    app.close()


def test_primary_rays_of_a_tiled_multi_device_launch(built, shade, tmp_path):
    """distribute(): device d of 3 renders the tiles (x + y) % 3 == d of a 100 x 40 image (a width that is not a multiple of the
    tile grid, so some launch indices fall outside the image and are skipped)."""
    app = host.App(H.write_system(tmp_path, "rtigo3_geometry", resolution="100 40", samplesSqrt=1), H.scene_path("rtigo3_geometry"), host_only=True)
    ref = H.oracle_scene(app)
    base = app.system_data(0)
    skipped = 0
    for device in range(3):
        camera = np.ascontiguousarray(app.camera())
        sysd = host.SystemData.from_buffer_copy(bytes(base))
        sysd.cameraDefinitions = camera.ctypes.data
        sysd.deviceCount, sysd.deviceIndex, sysd.distribution = 3, device, 1
        launch_width = ((100 + 8 * 3 - 1) // (8 * 3)) * 8          # tiles of 8 columns dealt to 3 devices
        rays = np.zeros(launch_width * 40, dtype=orc.RAY_DTYPE)
        shade.sh_generate_primary(C.byref(sysd), launch_width, 40, 0, rays.ctypes.data)
        osys = H.oracle_sys(app)
        osys.deviceCount, osys.deviceIndex, osys.distribution = 3, device, 1
        want = ref.generate_primary(osys, launch_width, 40, 0)
        assert rays_identical(rays, want)
        skipped += int((rays["tmax"] < 0).sum())
    assert skipped > 0
    app.close()


def test_texture_fetch_equals_the_oracle(built, shade):
    rng = np.random.default_rng(5)
    w, h = 13, 7
    block = np.zeros(4 + 4 * w * h, dtype=np.float32)
    block[:4].view(np.uint32)[:] = [w, h, 0, 0]
    block[4:] = rng.uniform(0.0, 1.0, size=4 * w * h).astype(np.float32)
    handle = block.ctypes.data
    uv = np.concatenate([rng.uniform(0, 1, size=(500, 2)), rng.uniform(-7, 9, size=(500, 2)),
                         [[0.0, 0.0], [1.0, 1.0], [0.5 / w, 0.5 / h], [1.0 - 1e-7, 1e-7]]]).astype(np.float32)
    got = np.zeros((len(uv), 3), dtype=np.float32)
    shade.sh_tex2d(handle, len(uv), uv.ctypes.data, got.ctypes.data)
    L = orc.lib()
    L.orc_tex2d.argtypes = [C.c_uint64, C.c_float, C.c_float, C.POINTER(C.c_float * 3)]
    L.orc_tex2d.restype = None
    want = np.zeros_like(got)
    for i, (u, v) in enumerate(uv):
        out = (C.c_float * 3)()
        L.orc_tex2d(handle, float(u), float(v), C.byref(out))
        want[i] = out[:]
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def _bind_callables(shade):
    L = orc.lib()
    for lib, prefix in ((shade, "sh_"), (L, "orc_test_")):
        getattr(lib, prefix + "bsdf_sample").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        getattr(lib, prefix + "bsdf_sample").restype = None
        getattr(lib, prefix + "bsdf_eval").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        getattr(lib, prefix + "bsdf_eval").restype = None
        getattr(lib, prefix + "light_constant").argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        getattr(lib, prefix + "light_constant").restype = None
        getattr(lib, prefix + "light_parallelogram").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        getattr(lib, prefix + "light_parallelogram").restype = None
    return L


def _unit(rng, n):
    v = rng.normal(size=(n, 3))
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("bsdf", [0, 1, 2, 3, 4])
def test_bsdf_callables_equal_the_oracle_word_for_word(built, shade, bsdf):
    """The five BSDF sample callables and the two non-trivial eval callables of csrc/shade.cuh (diffuse, specular reflection,
    specular transmission with nested-volume bookkeeping, GGX-Smith reflection and transmission) against the oracle's
    restatement of bxdf_diffuse.cu / bxdf_specular.cu / bxdf_ggx_smith.cu: all 28 words of the per-ray data after the call,
    random materials, frames, directions from both sides of the surface, flag combinations and seeds."""
    L = _bind_callables(shade)
    rng = np.random.default_rng(100 + bsdf)
    n = 4000
    normal, wo = _unit(rng, n), _unit(rng, n)
    tangent = np.cross(normal, _unit(rng, n)).astype(np.float32)
    tangent /= np.maximum(np.linalg.norm(tangent, axis=1, keepdims=True), 1e-6).astype(np.float32)
    flip = rng.random(n) < 0.3                         # geometric normal on the other side for some
    normal_geo = np.where(flip[:, None], -normal, normal).astype(np.float32)
    mismatches = transmitted = terminated = sampled = 0
    for i in range(n):
        m = np.zeros(1, dtype=orc.MATERIAL_DTYPE)
        m["indexBSDF"] = bsdf
        m["roughness"] = rng.uniform(0.02, 0.9, 2).astype(np.float32) if rng.random() < 0.8 else np.float32([0.0, 0.0])
        m["albedo"] = rng.uniform(0.05, 1.0, 3)
        m["absorption"] = rng.uniform(0.0, 2.0, 3)
        m["ior"] = rng.uniform(1.01, 2.4)
        thin = rng.random() < 0.25
        m["flags"] = 0x20 if thin else 0
        st = np.concatenate([normal_geo[i], tangent[i], normal[i], rng.uniform(0.05, 1.0, 3)]).astype(np.float32)
        prd = np.zeros(28, dtype=np.uint32)
        f = prd.view(np.float32)
        f[0:3] = rng.uniform(-2, 2, 3); f[3] = rng.uniform(0.1, 5.0)
        f[4:7] = wo[i]
        f[10:13] = rng.uniform(0, 1, 3)
        prd[13] = (0x1 | (0x10 if rng.random() < 0.6 else 0) | (0x20 if thin else 0) | (0x1000 if rng.random() < 0.2 else 0))
        f[14:17] = 1.0; f[17] = 1.0
        f[18:21] = rng.uniform(0, 1, 3)
        f[21:23] = [rng.uniform(1.0, 2.0), rng.uniform(1.0, 2.0)]
        f[23:27] = rng.uniform(0.0, 2.0, 4)
        prd[27] = rng.integers(0, 2 ** 32, dtype=np.uint64)
        a, b = prd.copy(), prd.copy()
        shade.sh_bsdf_sample(m.ctypes.data, st.ctypes.data, a.ctypes.data)
        L.orc_test_bsdf_sample(m.ctypes.data, st.ctypes.data, b.ctypes.data)
        if not np.array_equal(a, b):
            mismatches += 1
        transmitted += int(bool(a[13] & 0x100))
        terminated += int(bool(a[13] & 0x80000000))
        sampled += int(np.any(a[7:10] != 0))
        wi_l = _unit(rng, 1)[0]
        ea, eb = np.zeros(4, dtype=np.float32), np.zeros(4, dtype=np.float32)
        shade.sh_bsdf_eval(m.ctypes.data, st.ctypes.data, a.ctypes.data, wi_l.ctypes.data, ea.ctypes.data)
        L.orc_test_bsdf_eval(m.ctypes.data, st.ctypes.data, b.ctypes.data, wi_l.ctypes.data, eb.ctypes.data)
        if not np.array_equal(ea.view(np.uint32), eb.view(np.uint32)):
            mismatches += 1
    assert mismatches == 0
    # the cases are not degenerate: directions are sampled, transmissive lobes transmit, some samples end the path
    assert sampled > 0.5 * n
    assert (transmitted > 0.05 * n) == (bsdf in (2, 4))
    if bsdf in (0, 3):
        assert terminated > 0


def test_light_callables_equal_the_oracle(built, shade):
    L = _bind_callables(shade)
    rng = np.random.default_rng(9)
    for _ in range(3000):
        sample = rng.random(2).astype(np.float32)
        a, b = np.zeros(8, dtype=np.float32), np.zeros(8, dtype=np.float32)
        num = int(rng.integers(1, 4))
        shade.sh_light_constant(num, sample.ctypes.data, a.ctypes.data)
        L.orc_test_light_constant(num, sample.ctypes.data, b.ctypes.data)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        light = np.zeros(1, dtype=orc.LIGHT_DTYPE)
        light["type"] = 1
        light["position"] = rng.uniform(-2, 2, 3)
        u, v = rng.uniform(-2, 2, 3), rng.uniform(-2, 2, 3)
        light["vecU"], light["vecV"] = u, v
        nrm = np.cross(u, v)
        light["area"] = np.linalg.norm(nrm)
        light["normal"] = nrm / max(np.linalg.norm(nrm), 1e-6)
        light["emission"] = rng.uniform(0, 20, 3)
        point = rng.uniform(-3, 3, 3).astype(np.float32)
        shade.sh_light_parallelogram(light.ctypes.data, num, point.ctypes.data, sample.ctypes.data, a.ctypes.data)
        L.orc_test_light_parallelogram(light.ctypes.data, num, point.ctypes.data, sample.ctypes.data, b.ctypes.data)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_miss_programs_and_the_environment_light_equal_the_oracle(built, shade, tmp_path):
    """__miss__env_null / _constant / _sphere and the importance-sampled spherical environment light (CDF search, bilinear
    environment lookup with wrap in u and clamp in v, MIS weight) on the host's own environment map and CDF tables."""
    L = orc.lib()
    for lib, prefix in ((shade, "sh_"), (L, "orc_test_")):
        getattr(lib, prefix + "miss").argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_int, C.c_void_p]
        getattr(lib, prefix + "miss").restype = None
        getattr(lib, prefix + "light_sphere").argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
        getattr(lib, prefix + "light_sphere").restype = None
    app = host.App(H.write_system(tmp_path, "rtigo3_geometry", resolution="32 32", samplesSqrt=1, miss=2, envMap="procedural 64 32", envRotation=0.15),
                   H.scene_path("rtigo3_geometry"), host_only=True)
    texels, cdf_u, cdf_v, integral = app.environment()
    app.close()
    texels, cdf_u, cdf_v = (np.ascontiguousarray(a, dtype=np.float32) for a in (texels, cdf_u, cdf_v))
    h, w = texels.shape[:2]
    rng = np.random.default_rng(21)
    for i in range(3000):
        rotation = float(rng.random()) if i % 3 else 0.0
        sample = rng.random(2).astype(np.float32)
        a, b = np.zeros(8, dtype=np.float32), np.zeros(8, dtype=np.float32)
        num = int(rng.integers(1, 4))
        shade.sh_light_sphere(texels.ctypes.data, w, h, cdf_u.ctypes.data, cdf_v.ctypes.data, integral, rotation, num, sample.ctypes.data, a.ctypes.data)
        L.orc_test_light_sphere(texels.ctypes.data, w, h, cdf_u.ctypes.data, cdf_v.ctypes.data, integral, rotation, num, sample.ctypes.data, b.ctypes.data)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        assert a[7] > 0.0 and abs(float(np.linalg.norm(a[:3])) - 1.0) < 1e-4
        prd = np.zeros(28, dtype=np.uint32)
        f = prd.view(np.float32)
        f[7:10] = _unit(rng, 1)[0]
        if i % 50 == 0:
            f[7:10] = [0.0, 1.0 if i % 100 else -1.0, 0.0]          # the poles: clamp in v
        prd[13] = 0x1 | (0x4 if rng.random() < 0.5 else 0)         # RT_FLAG_DIFFUSE selects the MIS weight
        f[17] = rng.uniform(0.01, 3.0)
        for miss in (0, 1, 2):
            pa, pb = prd.copy(), prd.copy()
            shade.sh_miss(texels.ctypes.data, w, h, integral, rotation, miss, pa.ctypes.data)
            L.orc_test_miss(texels.ctypes.data, w, h, integral, rotation, miss, pb.ctypes.data)
            assert np.array_equal(pa, pb) and (pa[13] & 0x80000000)
