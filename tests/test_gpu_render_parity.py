"""GPU parity, part 2: the integrator.  Frames rendered by the wavefront kernels through the reference-facing
Application -> Raytracer -> Device classes must equal the scalar oracle's frames at identical TEA/LCG seeds.

Tolerance: NONE for the in-scope configurations.  The shading kernels are compiled without FMA contraction and take
their transcendentals from include/rt_portable_math.h, exactly like the oracle, so float32 radiance is compared as raw
bits.  (Should a future change make that impossible, the per-pixel bound to fall back to is |a-b| <= 1e-5 * max(1,|b|).)"""
import ctypes as C

import numpy as np
import pytest

import helpers as H
from oracle import orc
from tweeker_raytracer_b200 import core, host

pytestmark = pytest.mark.gpu


def render_both(tmp, name, iterations, batch=1, **overrides):
    app = host.App(H.write_system(tmp, name, **overrides), H.scene_path(name))
    try:
        ref = H.oracle_scene(app)
        done = 0
        while done < iterations:
            done = app.render(min(batch, iterations - done))
        got = app.frame()
        w, h = app.resolution
        st = orc.Stats()
        want = ref.render(H.oracle_sys(app), app.info.miss, w, h, iter_count=iterations, stats=st).reshape(h, w, 4)
        stats = app.stats()
        return got, want, stats, st
    finally:
        app.close()


def assert_frames_identical(got, want):
    same = got.view(np.uint32) == want.view(np.uint32)
    if not same.all():
        bad = np.argwhere(~same.all(axis=2))
        y, x = bad[0]
        raise AssertionError("%d of %d pixels differ; first at (x=%d, y=%d): gpu %r oracle %r, max abs diff %g"
                             % (len(bad), got.shape[0] * got.shape[1], x, y, got[y, x], want[y, x], np.nanmax(np.abs(got - want))))


def test_cornell_box_16spp_bit_exact(cuda_device, tmp_path):
    got, want, stats, st = render_both(tmp_path, "rtigo3_cornell_box", 16, resolution="128 128")
    assert_frames_identical(got, want)
    assert stats.pathSamples == st.pathSamples == 128 * 128 * 16
    assert stats.radianceRays == st.radianceRays and stats.shadowRays == st.shadowRays
    assert stats.kernelLaunches > 0


def test_cornell_box_batched_iterations_equal_single(cuda_device, tmp_path):
    a, want, _, _ = render_both(tmp_path, "rtigo3_cornell_box", 9, batch=4, resolution="96 64", samplesSqrt=3)
    assert_frames_identical(a, want)


def test_geometry_scene_all_bsdfs_bit_exact(cuda_device, tmp_path):
    got, want, stats, st = render_both(tmp_path, "rtigo3_geometry", 8, batch=8, resolution="240 136", samplesSqrt=3)
    assert_frames_identical(got, want)
    assert stats.radianceRays == st.radianceRays and stats.shadowRays == st.shadowRays


def test_geometry_scene_hdr_environment_bit_exact(cuda_device, tmp_path):
    got, want, _, _ = render_both(tmp_path, "rtigo3_geometry", 4, batch=2, resolution="200 112", samplesSqrt=2, miss=2,
                                  envMap="procedural 256 128", envRotation=0.15)
    assert_frames_identical(got, want)
    assert np.isfinite(got).all() and got[..., :3].mean() > 0.05


@pytest.mark.parametrize("lens", [1, 2])
def test_lens_shaders_bit_exact(cuda_device, tmp_path, lens):
    got, want, _, _ = render_both(tmp_path, "rtigo3_cornell_box", 2, resolution="96 96", samplesSqrt=2, lensShader=lens)
    assert_frames_identical(got, want)


def test_no_lights_black_environment(cuda_device, tmp_path):
    # miss 0 + light 0: numLights stays 0, NEE is skipped, every path ends black (Device.cpp:992-996, closesthit.cu:253)
    got, want, stats, _ = render_both(tmp_path, "rtigo3_cornell_box", 2, resolution="64 64", samplesSqrt=2, light=0)
    assert_frames_identical(got, want)
    assert stats.shadowRays == 0 and float(got[..., :3].max()) == 0.0


def test_white_furnace_converges_to_one(cuda_device, tmp_path):
    # one convex diffuse sphere of albedo 1 under the constant white environment, paths of exactly two segments:
    # NEE and the implicit environment hit are MIS-complementary, so every pixel converges to 1.
    scene = tmp_path / "scene_furnace.txt"
    scene.write_text("albedo 1 1 1\nmaterial default brdf_diffuse\nmaterial white brdf_diffuse\nidentity\nmodel sphere 64 32 1.0 white\n")
    sysfile = H.write_system(tmp_path, "rtigo3_cornell_box", resolution="64 64", samplesSqrt=16, miss=1, light=0,
                             pathLengths="2 2", center="0 0 0", camera="0.75 0.5 45 4")
    app = host.App(sysfile, str(scene))
    try:
        while app.render(64) < app.spp:
            pass
        img = app.frame()[..., :3]
    finally:
        app.close()
    assert abs(float(img.mean()) - 1.0) < 0.01
    assert float(np.abs(img.reshape(-1, 3).mean(axis=1) - 1.0).max()) < 0.35    # 256 spp noise bound


def emulate_devices(app, count, iterations, composite=True):
    """One GPU plays `count` devices of the tiled strategies: same launches the local-copy devices would enqueue."""
    ctx = app.context(0)
    w, h = app.resolution
    base = app.system_data(0)
    tile = base.tileSize.x
    lw = ((w + count - 1) // count + tile - 1) & ~(tile - 1)
    slabs = []
    for index in range(count):
        sys = host.SystemData.from_buffer_copy(bytes(base))
        sys.deviceCount, sys.deviceIndex, sys.distribution = count, index, 1
        sys.texelBuffer = ctx.malloc(lw * h * 16)
        ctx.memset(sys.texelBuffer, 0, lw * h * 16)
        ctx.launch(sys, lw, h, core.RAYGEN_LOCAL_COPY, app.info.miss, 0, iterations)
        slabs.append((sys, sys.texelBuffer))
    frame = ctx.malloc(w * h * 16)
    ctx.memset(frame, 0, w * h * 16)
    for index, (sys, slab) in enumerate(slabs):
        args = host.CompositorData()
        args.outputBuffer, args.tileBuffer = frame, slab
        args.resolution, args.tileSize, args.tileShift = sys.resolution, sys.tileSize, sys.tileShift
        args.launchWidth, args.deviceCount, args.deviceIndex = lw, count, index
        ctx.composite(args)
    out = ctx.download(frame, np.float32, w * h * 4).reshape(h, w, 4)
    texels = [ctx.download(slab, np.float32, lw * h * 4).reshape(h, lw, 4) for _, slab in slabs]
    for _, slab in slabs:
        ctx.free(slab)
    ctx.free(frame)
    return out, texels, lw


@pytest.mark.parametrize("count", [2, 3, 8])
def test_tile_distribution_and_compositor_bit_exact(cuda_device, tmp_path, count):
    # strategy 3 semantics (raygeneration.cu:152-164, :259-344; compositor.cu:38-65) with the device layout dependent seeds
    app = host.App(H.write_system(tmp_path, "rtigo3_cornell_box", resolution="100 48", samplesSqrt=2, tileSize="8 8"),
                   H.scene_path("rtigo3_cornell_box"))
    try:
        ref = H.oracle_scene(app)
        got, texels, lw = emulate_devices(app, count, 3)
        w, h = app.resolution
        want = np.zeros((h, w, 4), dtype=np.float32)
        for index in range(count):
            sys = H.oracle_sys(app)
            sys.deviceCount, sys.deviceIndex, sys.distribution = count, index, 1
            slab = ref.render(sys, app.info.miss, lw, h, local_copy=True, iter_count=3).reshape(h, lw, 4)
            assert texels[index].tobytes() == slab.tobytes()
            args = orc.CompositorData()
            args.resolution.x, args.resolution.y = w, h
            args.tileSize.x, args.tileSize.y = sys.tileSize.x, sys.tileSize.y
            args.tileShift.x, args.tileShift.y = sys.tileShift.x, sys.tileShift.y
            args.launchWidth, args.deviceCount, args.deviceIndex = lw, count, index
            orc.composite(args, slab, want)
        assert got.tobytes() == want.tobytes()
        assert (got[..., 3] == 1.0).all()      # every pixel is owned by exactly one device
    finally:
        app.close()


def test_tonemap_kernel_matches_oracle(cuda_device, tmp_path):
    app = host.App(H.write_system(tmp_path, "rtigo3_cornell_box", resolution="128 128"), H.scene_path("rtigo3_cornell_box"))
    try:
        app.render(4)
        frame = app.frame()
        got = app.tonemap()
        t = app.tonemapper()
        params = orc.TonemapperParams(t.gamma, tuple(t.colorBalance), t.whitePoint, t.burnHighlights, t.crushBlacks, t.saturation, t.brightness)
        want = orc.tonemap(params, frame).reshape(got.shape)
        assert np.array_equal(got, want)
        assert got.max() > 200 and got.min() == 0
    finally:
        app.close()


def test_update_restarts_accumulation(cuda_device, tmp_path):
    app = host.App(H.write_system(tmp_path, "rtigo3_cornell_box", resolution="64 64", samplesSqrt=2), H.scene_path("rtigo3_cornell_box"))
    try:
        assert app.render(3) == 3
        first = app.frame().copy()
        assert app.render(8) == 4          # clamped to samplesSqrt^2
        app.restart()
        assert app.render(3) == 3
        assert app.frame().tobytes() == first.tobytes()
    finally:
        app.close()


def test_mutators_restart_and_match_oracle(cuda_device, tmp_path):
    # updateCamera / updateMaterial / updateLight (Raytracer.cpp:331-367): each restarts the accumulation, and the frame
    # rendered afterwards equals the oracle's frame of the modified scene
    app = host.App(H.write_system(tmp_path, "rtigo3_cornell_box", resolution="96 96", samplesSqrt=2), H.scene_path("rtigo3_cornell_box"))
    try:
        app.render(4)
        app.set_camera(0.7, 0.55, 50.0, 3.2, (0.0, 1.0, 0.0))
        app.update_material(5, 3, (0.9, 0.6, 0.2), roughness=(0.3, 0.1))                                   # mirror sphere -> anisotropic GGX
        app.update_material(6, 2, (1, 1, 1), absorption_color=(0.5, 0.8, 0.9), absorption_scale=2.0, ior=1.33)   # glass with absorption
        app.update_light_emission(0, (4.0, 8.0, 12.0))
        assert app.render(3) == 3
        got = app.frame()
        ref = H.oracle_scene(app)
        want = ref.render(H.oracle_sys(app), app.info.miss, 96, 96, iter_count=3).reshape(96, 96, 4)
        assert_frames_identical(got, want)
        assert app.materials()["indexBSDF"][5] == 3 and abs(app.materials()["ior"][6] - 1.33) < 1e-6
        assert app.stats().stackOverflows == 0
    finally:
        app.close()


@pytest.mark.parametrize("key", ["geometry_converged_96x54_256spp", "cornell_converged_64x64_256spp"])
def test_converged_frames_reach_40_db_against_the_reference(cuda_device, tmp_path, key):
    """BASELINE.json: converged images at PSNR >= 40 dB against the reference's output -- here the committed frames of the
    reference's own device code compiled for the host (tests/golden/make_golden.py), rendered on the GPU with the same seeds."""
    import os
    from golden import make_golden
    golden = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_frames.npz"))
    name, overrides, iterations = make_golden.CASES[key]
    with host.App(H.write_system(tmp_path, name, **overrides), H.scene_path(name)) as app:
        while app.render(64) < iterations:
            pass
        got = app.frame()
    want = golden[key]
    a, b = np.clip(got[..., :3], 0.0, 1.0), np.clip(want[..., :3], 0.0, 1.0)
    assert H.psnr(a, b) >= 40.0, H.psnr(a, b)


def test_coalesced_render_calls_equal_single_launches(cuda_device, tmp_path):
    """The reference's calling pattern -- `unsigned int render()` once per iteration (Application.cpp:500-503) -- is coalesced
    into batched launches behind the unchanged signature; the frame must not depend on where the batches fall."""
    frames = []
    for limit in (1, 4, 0):          # every call launches / batches of four / the default limit
        app = host.App(H.write_system(tmp_path, "rtigo3_cornell_box", resolution="96 64", samplesSqrt=3), H.scene_path("rtigo3_cornell_box"))
        try:
            if limit:
                app.set_coalesce(limit)
            assert app.render_calls(5) == 5            # iterations "done" are counted at once, launched lazily
            before = app.stats().kernelLaunches         # stats() is an observation point: it flushes
            assert before > 0
            assert app.render_calls(4) == 9
            assert app.render_calls(3) == 9            # the budget is samplesSqrt^2
            frames.append(app.frame())
            if not limit:
                ref = H.oracle_scene(app)
                w, h = app.resolution
                want = ref.render(H.oracle_sys(app), app.info.miss, w, h, iter_count=9).reshape(h, w, 4)
        finally:
            app.close()
    assert_frames_identical(frames[0], want)
    assert_frames_identical(frames[1], want)
    assert_frames_identical(frames[2], want)


def test_geometry_scene_hdr_file_environment_bit_exact(cuda_device, tmp_path):
    """`miss 2` with the environment read from a Radiance .hdr FILE (reference: Texture.cpp:1300-1377 through
    DevIL; here host/ImageIO.cpp), importance sampled through the CDFs the host builds from it (pinned against the reference's
    own builder in test_cpu_host_loader.py), rendered on the GPU and compared with the oracle bit for bit."""
    path = H.write_rgbe_hdr(str(tmp_path / "sky.hdr"), H.random_rgbe(64, 32, 7))
    got, want, stats, st = render_both(tmp_path, "rtigo3_geometry", 4, batch=4, resolution="160 90", samplesSqrt=2, miss=2,
                                       envMap=path, envRotation=0.35)
    assert_frames_identical(got, want)
    assert stats.radianceRays == st.radianceRays and stats.shadowRays == st.shadowRays
    assert np.isfinite(got).all() and got[..., :3].mean() > 0.01
