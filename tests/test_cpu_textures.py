"""Material textures on the CPU side (no GPU): picture loading, the texture fetch the repository defines, and properties of
the ordered any-hit processing of the oracle (the pin against the reference's own any-hit programs is in
test_cpu_oracle_vs_reference.py, cases textures_*)."""
import os
import struct
import zlib

import numpy as np

import helpers as H
from oracle import orc
from tweeker_raytracer_b200 import host


def _write_png(path, rgba8, colour=6, filters=(0, 1, 2, 3, 4)):
    """Minimal PNG writer exercising all five scanline filters (so the reader's unfilter code is covered)."""
    h, w, c = rgba8.shape
    raw = bytearray()
    prev = np.zeros((w, c), dtype=np.int32)
    for y in range(h):
        f = filters[y % len(filters)]
        line = rgba8[y].astype(np.int32)
        left = np.vstack([np.zeros((1, c), dtype=np.int32), line[:-1]])
        upleft = np.vstack([np.zeros((1, c), dtype=np.int32), prev[:-1]])
        if f == 0:
            out = line
        elif f == 1:
            out = line - left
        elif f == 2:
            out = line - prev
        elif f == 3:
            out = line - (left + prev) // 2
        else:
            p = left + prev - upleft
            pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - upleft)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, upleft))
            out = line - pred
        raw.append(f)
        raw += (out & 255).astype(np.uint8).tobytes()
        prev = line

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    with open(path, "wb") as fh:
        fh.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, colour, 0, 0, 0)) +
                 chunk(b"IDAT", zlib.compress(bytes(raw))) + chunk(b"IEND", b""))


def _app(tmp_path, **kw):
    opts = dict(resolution="48 27", samplesSqrt=2)
    opts.update(kw)
    return host.App(H.write_system(tmp_path, "rtigo3_textures", **opts), H.scene_path("rtigo3_textures"), host_only=True)


def test_procedural_pictures_when_the_files_are_absent(built, tmp_path):
    with _app(tmp_path) as app:
        albedo, cutout = app.picture("albedo"), app.picture("cutout")
        assert albedo.shape == (256, 256, 4) and cutout.shape == (256, 256, 4)
        assert set(np.unique(cutout[..., 0])) == {0.0, 0.5, 1.0}          # holes, the half-transparent band, bars
        assert np.all(albedo[..., 3] == 1.0) and albedo[..., :3].max() <= 1.0


def test_png_and_ppm_pictures(built, tmp_path):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, size=(13, 17, 4), dtype=np.uint8)
    png = os.path.join(str(tmp_path), "albedo.png")
    _write_png(png, img)
    ppm = os.path.join(str(tmp_path), "cutout.ppm")
    with open(ppm, "wb") as f:
        f.write(b"P6\n# comment\n17 13\n255\n" + img[..., :3].tobytes())
    with _app(tmp_path, textureAlbedo=png, textureCutout=ppm) as app:
        a, c = app.picture("albedo"), app.picture("cutout")
        want = img[::-1].astype(np.float32) / np.float32(255.0)           # texel row 0 = bottom row of the file
        assert np.array_equal(a, want)
        assert np.array_equal(c[..., :3], want[..., :3]) and np.all(c[..., 3] == 1.0)
    grey = os.path.join(str(tmp_path), "grey.png")
    _write_png(grey, img[..., :1], colour=0)
    with _app(tmp_path, textureCutout=grey) as app:
        c = app.picture("cutout")
        assert np.array_equal(c[..., 0], img[::-1, :, 0].astype(np.float32) / np.float32(255.0)) and np.array_equal(c[..., 0], c[..., 2])


def test_texture_fetch_definition(built, tmp_path):
    """Bilinear, texel centres at (i + 0.5) / size, wrap in u and v."""
    with _app(tmp_path) as app:
        m = app.materials()
        handle = int(m["textureCutout"][m["textureCutout"] != 0][0])
        tex = app.picture("cutout")
        L = orc.lib()
        import ctypes as C
        L.orc_tex2d.argtypes = [C.c_uint64, C.c_float, C.c_float, C.POINTER(C.c_float * 3)]
        L.orc_tex2d.restype = None

        def fetch(u, v):
            out = (C.c_float * 3)()
            L.orc_tex2d(handle, u, v, C.byref(out))
            return np.array(out[:], dtype=np.float32)
        H_, W_ = tex.shape[:2]
        for (x, y) in [(0, 0), (5, 20), (255, 255), (100, 63)]:
            assert np.array_equal(fetch((x + 0.5) / W_, (y + 0.5) / H_), tex[y, x, :3])               # texel centre = the texel
            assert np.array_equal(fetch((x + 0.5) / W_ + 3.0, (y + 0.5) / H_ - 2.0), tex[y, x, :3])   # wraps both ways
        # halfway between two texels of different value
        y0 = 15                                      # band 0 (1.0) -> band 1 at row 16
        x = 10                                       # inside a slot: band 1 there is 0.0
        mid = fetch((x + 0.5) / W_, (y0 + 1.0) / H_)
        assert np.allclose(mid, 0.5 * (tex[y0, x, :3] + tex[y0 + 1, x, :3]))
        # the left edge blends with the right edge (wrap)
        edge = fetch(0.0, (20 + 0.5) / H_)
        assert np.allclose(edge, 0.5 * (tex[20, 0, :3] + tex[20, W_ - 1, :3]))


def test_candidates_are_enumerated_in_canonical_order(built, tmp_path):
    """orc_trace_closest_after walks the hits of a ray in ascending (t, instance, primitive) order and visits every
    intersection the brute-force intersector knows."""
    import ctypes as C
    with _app(tmp_path) as app:
        scene = H.oracle_scene(app)
        L = orc.lib()
        L.orc_trace_closest_after.restype = None
        rays = H.random_rays(300, seed=11, lo=(-6, 0.2, -4), hi=(6, 3, 4))
        hit = np.zeros(1, dtype=orc.HIT_DTYPE)
        for r in range(len(rays)):
            ray = rays[r:r + 1]
            seq = []
            first = scene.trace_closest(ray)[0]
            cur = first
            while cur["inst"] != 0xffffffff and len(seq) < 64:
                seq.append((float(cur["t"]), int(cur["inst"]), int(cur["prim"])))
                L.orc_trace_closest_after(scene.h, ray.ctypes.data_as(C.c_void_p), C.c_float(cur["t"]), C.c_uint32(int(cur["inst"])),
                                          C.c_uint32(int(cur["prim"])), hit.ctypes.data_as(C.c_void_p))
                cur = hit[0].copy()
            assert seq == sorted(seq)
            assert len(set(seq)) == len(seq)
            # shrinking tmax to just above the k-th candidate makes it the last one: the enumeration misses nothing
            if len(seq) >= 2:
                cut = ray.copy()
                cut["tmax"] = np.nextafter(np.float32(seq[1][0]), np.float32(np.inf))
                again = scene.trace_closest(cut, brute_force=True)[0]
                assert (float(again["t"]), int(again["inst"]), int(again["prim"])) == seq[0]


def test_opaque_cutout_draws_nothing(built, tmp_path):
    """opacity == 1 short-circuits the random draw (anyhit.cu:75, :122): a fully opaque cutout picture leaves the image unchanged."""
    ones = os.path.join(str(tmp_path), "ones.ppm")
    with open(ones, "wb") as f:
        f.write(b"P5\n2 2\n255\n" + bytes([255] * 4))
    with _app(tmp_path, textureCutout=ones) as app:
        a = H.oracle_scene(app).render(H.oracle_sys(app), app.info.miss, 48, 27, iter_count=3)
        m = app.materials()
        m["textureCutout"] = 0
        s = H.oracle_scene(app)
        s.set_materials(m)
        b = s.render(H.oracle_sys(app), app.info.miss, 48, 27, iter_count=3)
        assert a.tobytes() == b.tobytes()


def test_cutout_changes_the_image_and_lets_light_through(built, tmp_path):
    with _app(tmp_path, resolution="64 36") as app:
        sysd = H.oracle_sys(app)
        textured = H.oracle_scene(app).render(sysd, app.info.miss, 64, 36, iter_count=4)
        m = app.materials()
        m["textureCutout"] = 0
        m["textureAlbedo"] = 0
        s = H.oracle_scene(app)
        s.set_materials(m)
        plain = s.render(sysd, app.info.miss, 64, 36, iter_count=4)
        assert textured.tobytes() != plain.tobytes()
        assert np.isfinite(textured).all() and textured[..., :3].min() >= 0.0


def test_candidate_order_with_exact_ties(built, tmp_path):
    """Two coincident planes (same transform, different instances) and a ray through their shared diagonal: four candidates at
    the SAME t.  The canonical order resolves them by (instance, primitive), and the enumeration visits each exactly once."""
    import ctypes as C
    scene_file = os.path.join(str(tmp_path), "scene_ties.txt")
    with open(scene_file, "w") as f:
        f.write("albedo 1 1 1\ncutoutTexture 1\nmaterial a brdf_diffuse\nmaterial b brdf_diffuse\nidentity\n"
                "push translate 0 1 0 model plane 1 1 1 a pop\npush translate 0 1 0 model plane 1 1 1 b pop\n"
                "push translate 0 2 0 model plane 1 1 1 a pop\n")
    with host.App(H.write_system(tmp_path, "rtigo3_textures", resolution="8 8", light=0, miss=1), scene_file, host_only=True) as app:
        scene = H.oracle_scene(app)
        L = orc.lib()
        L.orc_trace_closest_after.restype = None
        ray = np.zeros(1, dtype=orc.RAY_DTYPE)
        ray["ox"], ray["oy"], ray["oz"] = 0.0, 0.0, 0.0          # straight up through the centre: on the diagonal both triangles share
        ray["dx"], ray["dy"], ray["dz"] = 0.0, 1.0, 0.0
        ray["tmin"], ray["tmax"] = 1e-4, 1e27
        hit = np.zeros(1, dtype=orc.HIT_DTYPE)
        seq, cur = [], scene.trace_closest(ray)[0]
        while cur["inst"] != 0xffffffff and len(seq) < 16:
            seq.append((float(cur["t"]), int(cur["inst"]), int(cur["prim"])))
            L.orc_trace_closest_after(scene.h, ray.ctypes.data_as(C.c_void_p), C.c_float(cur["t"]), C.c_uint32(int(cur["inst"])),
                                      C.c_uint32(int(cur["prim"])), hit.ctypes.data_as(C.c_void_p))
            cur = hit[0].copy()
        assert seq == sorted(seq) and len(set(seq)) == len(seq)
        at_one = [s for s in seq if s[0] == 1.0]
        assert [(i, p) for _, i, p in at_one] == [(0, 0), (0, 1), (1, 0), (1, 1)]     # both triangles of both planes, tie-ordered
        assert [s[1] for s in seq if s[0] == 2.0] == [2, 2]
        # the frame renders (the stochastic test runs on tied candidates without looping)
        img = scene.render(H.oracle_sys(app), app.info.miss, 8, 8, iter_count=2)
        assert np.isfinite(img).all()


def test_damaged_picture_files_fall_back_to_the_procedural_pictures(built, tmp_path):
    """A truncated PNG, a PNG with a wrong signature and an interlaced PNG are rejected by the reader; the Application then uses
    the procedural picture (the reference would assert)."""
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(9, 7, 4), dtype=np.uint8)
    good = os.path.join(str(tmp_path), "good.png")
    _write_png(good, img)
    data = open(good, "rb").read()
    cases = {"truncated.png": data[: len(data) // 2], "signature.png": b"\x89PNX" + data[4:], "empty.png": b""}
    interlaced = bytearray(data)
    interlaced[28] = 1                                   # IHDR interlace method (offset 8 + 8 + 12)
    cases["interlaced.png"] = bytes(interlaced)          # (the CRC no longer matches; the reader rejects the method first)
    for name, blob in cases.items():
        path = os.path.join(str(tmp_path), name)
        with open(path, "wb") as f:
            f.write(blob)
        with _app(tmp_path, textureAlbedo=path) as app:
            assert app.picture("albedo").shape == (256, 256, 4), name      # procedural fallback
    with _app(tmp_path, textureAlbedo=good) as app:
        assert app.picture("albedo").shape == (9, 7, 4)
