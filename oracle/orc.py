"""ctypes binding of the scalar CPU oracle (oracle/rt_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module;
nothing under tweeker_raytracer_b200/ does.  See oracle/rt_oracle.h for what each function restates.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIBS = {"pinned": os.path.join(_HERE, "_build", "liborc.so"), "libm": os.path.join(_HERE, "_build", "liborc_libm.so")}
REF_LIB = os.path.join(_HERE, "_ref", "libref.so")
REFERENCE_SHADERS = "/root/reference/apps/rtigo3/shaders"


def build(variant="pinned", force=False):
    """Compiles the oracle with gcc (seconds).  variant "pinned": transcendentals of include/rt_portable_math.h (what the
    GPU kernels are compared with); "libm": libm transcendentals (what the host-compiled reference is compared with)."""
    out = _LIBS[variant]
    src = os.path.join(_HERE, "rt_oracle.c")
    deps = [src, os.path.join(_HERE, "rt_oracle.h"), os.path.join(_ROOT, "include", "rtigo3_abi.h"),
            os.path.join(_ROOT, "include", "rt_portable_math.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-I" + os.path.join(_ROOT, "include"), "-I" + _HERE]
    if variant == "libm":
        cmd.append("-DRT_MATH_LIBM")
    subprocess.check_call(cmd + ["-o", out, src, "-lm", "-lpthread"])
    return out


def build_reference():
    """Compiles the reference's shader sources for the host (oracle/Makefile target `ref`).  Needs /root/reference; on a
    machine without it the prebuilt oracle/_ref/libref.so is used as is.  Returns the library path or None."""
    if os.path.isdir(REFERENCE_SHADERS):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
    return REF_LIB if os.path.exists(REF_LIB) else None


def reference_available():
    return os.path.exists(REF_LIB) or os.path.isdir(REFERENCE_SHADERS)


class Float3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Int2(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int)]


class SystemData(C.Structure):
    """rt_SystemData (include/rtigo3_abi.h), 192 bytes."""
    _fields_ = [("rect", C.c_int * 4), ("topObject", C.c_uint64), ("outputBuffer", C.c_uint64), ("tileBuffer", C.c_uint64),
                ("texelBuffer", C.c_uint64), ("cameraDefinitions", C.c_uint64), ("lightDefinitions", C.c_uint64),
                ("materialDefinitions", C.c_uint64), ("envTexture", C.c_uint64), ("envCDF_U", C.c_uint64), ("envCDF_V", C.c_uint64),
                ("resolution", Int2), ("tileSize", Int2), ("tileShift", Int2), ("pathLengths", Int2),
                ("deviceCount", C.c_int), ("deviceIndex", C.c_int), ("distribution", C.c_int), ("iterationIndex", C.c_int),
                ("samplesSqrt", C.c_int), ("sceneEpsilon", C.c_float), ("clockScale", C.c_float), ("lensShader", C.c_int),
                ("numCameras", C.c_int), ("numMaterials", C.c_int), ("numLights", C.c_int), ("envWidth", C.c_uint),
                ("envHeight", C.c_uint), ("envIntegral", C.c_float), ("envRotation", C.c_float), ("_pad", C.c_int)]


assert C.sizeof(SystemData) == 192


class CompositorData(C.Structure):
    _fields_ = [("outputBuffer", C.c_uint64), ("tileBuffer", C.c_uint64), ("resolution", Int2), ("tileSize", Int2),
                ("tileShift", Int2), ("launchWidth", C.c_int), ("deviceCount", C.c_int), ("deviceIndex", C.c_int), ("_pad", C.c_int)]


assert C.sizeof(CompositorData) == 56


class TonemapperParams(C.Structure):
    _fields_ = [("gamma", C.c_float), ("colorBalance", C.c_float * 3), ("whitePoint", C.c_float), ("burnHighlights", C.c_float),
                ("crushBlacks", C.c_float), ("saturation", C.c_float), ("brightness", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("radianceRays", C.c_uint64), ("shadowRays", C.c_uint64), ("pathSamples", C.c_uint64),
                ("nodesVisited", C.c_uint64), ("trisTested", C.c_uint64), ("instancesEntered", C.c_uint64)]


RAY_DTYPE = np.dtype([("ox", "f4"), ("oy", "f4"), ("oz", "f4"), ("tmin", "f4"), ("dx", "f4"), ("dy", "f4"), ("dz", "f4"), ("tmax", "f4")])
HIT_DTYPE = np.dtype([("t", "f4"), ("u", "f4"), ("v", "f4"), ("inst", "u4"), ("prim", "u4")])
ATTR_DTYPE = np.dtype([("vertex", "f4", 3), ("tangent", "f4", 3), ("normal", "f4", 3), ("texcoord", "f4", 3)])
MATERIAL_DTYPE = np.dtype([("textureAlbedo", "u8"), ("textureCutout", "u8"), ("roughness", "f4", 2), ("indexBSDF", "i4"),
                           ("albedo", "f4", 3), ("absorption", "f4", 3), ("ior", "f4"), ("flags", "u4"), ("pad0", "i4")])
LIGHT_DTYPE = np.dtype([("type", "i4"), ("position", "f4", 3), ("vecU", "f4", 3), ("vecV", "f4", 3), ("normal", "f4", 3),
                        ("area", "f4"), ("emission", "f4", 3), ("unused", "f4", 3)])
CAMERA_DTYPE = np.dtype([("P", "f4", 3), ("U", "f4", 3), ("V", "f4", 3), ("W", "f4", 3)])
assert ATTR_DTYPE.itemsize == 48 and MATERIAL_DTYPE.itemsize == 64 and LIGHT_DTYPE.itemsize == 80 and CAMERA_DTYPE.itemsize == 48

_loaded = {}


def lib(variant="pinned"):
    if variant not in _loaded:
        build(variant)
        L = C.CDLL(_LIBS[variant], mode=C.RTLD_GLOBAL if variant == "libm" else C.RTLD_LOCAL)
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_scene_add_geometry.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.orc_scene_add_instance.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_scene_set_materials.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_scene_set_lights.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_scene_set_camera.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_scene_set_env.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_float]
        L.orc_scene_commit.argtypes = [C.c_void_p]
        L.orc_scene_get_inverse.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_trace_any.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_generate_primary.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
        L.orc_render.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_void_p, C.c_void_p]
        L.orc_path_radiance.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_composite.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_tonemap.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_tea4.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_tea4.restype = C.c_uint32
        L.orc_rng.argtypes = [C.POINTER(C.c_uint32)]
        L.orc_rng.restype = C.c_float
        L.orc_online_cores.restype = C.c_int
        L.orc_uses_libm.restype = C.c_int
        _loaded[variant] = L
    return _loaded[variant]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Scene:
    """One oracle scene: geometries, instances, materials, lights, camera, optional environment."""

    def __init__(self, variant="pinned"):
        self.L = lib(variant)
        self.variant = variant
        self.h = C.c_void_p(self.L.orc_scene_create())
        self.num_instances = 0
        self.keep = {"geometries": [], "instances": []}   # host arrays the reference driver points into

    def close(self):
        if self.h:
            self.L.orc_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_geometry(self, attrs, indices):
        attrs = np.ascontiguousarray(attrs, dtype=ATTR_DTYPE)
        indices = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        self.keep["geometries"].append((attrs, indices))
        return self.L.orc_scene_add_geometry(self.h, _ptr(attrs), len(attrs), _ptr(indices), len(indices) // 3)

    def add_instance(self, transform, geometry, material, light=-1):
        t = np.ascontiguousarray(transform, dtype=np.float32).reshape(12)
        self.num_instances += 1
        self.keep["instances"].append((t.copy(), int(geometry), int(material), int(light)))
        return self.L.orc_scene_add_instance(self.h, _ptr(t), int(geometry), int(material), int(light))

    def set_materials(self, materials):
        m = np.ascontiguousarray(materials, dtype=MATERIAL_DTYPE)
        self.keep["materials"] = m
        self.L.orc_scene_set_materials(self.h, _ptr(m), len(m))

    def set_lights(self, lights):
        l = np.ascontiguousarray(lights, dtype=LIGHT_DTYPE)
        self.keep["lights"] = l
        self.L.orc_scene_set_lights(self.h, _ptr(l), len(l))

    def set_camera(self, camera):
        c = np.ascontiguousarray(camera, dtype=CAMERA_DTYPE).reshape(1)
        self.keep["camera"] = c
        self.L.orc_scene_set_camera(self.h, _ptr(c))

    def set_env(self, rgba, cdf_u, cdf_v, integral):
        rgba = np.ascontiguousarray(rgba, dtype=np.float32)
        h, w = rgba.shape[0], rgba.shape[1]
        cu = np.ascontiguousarray(cdf_u, dtype=np.float32)
        cv = np.ascontiguousarray(cdf_v, dtype=np.float32)
        self.keep["env"] = (rgba, cu, cv, float(integral))
        self.L.orc_scene_set_env(self.h, _ptr(rgba), w, h, _ptr(cu), _ptr(cv), C.c_float(integral))

    def commit(self):
        self.L.orc_scene_commit(self.h)

    def inverse(self, instance):
        out = np.zeros(12, dtype=np.float32)
        self.L.orc_scene_get_inverse(self.h, int(instance), _ptr(out))
        return out

    def trace_closest(self, rays, brute_force=False, stats=None):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        self.L.orc_trace_closest(self.h, _ptr(rays), len(rays), 1 if brute_force else 0, _ptr(hits), C.byref(stats) if stats is not None else None)
        return hits

    def trace_any(self, rays, brute_force=False, stats=None):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        occ = np.zeros(len(rays), dtype=np.uint8)
        self.L.orc_trace_any(self.h, _ptr(rays), len(rays), 1 if brute_force else 0, _ptr(occ), C.byref(stats) if stats is not None else None)
        return occ

    def trace_closest_after(self, rays, keys):
        """Closest candidate AFTER the key (t bits, instance, primitive) in the canonical order, per ray (orc_trace_closest_after:
        the step of the ordered any-hit processing)."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        keys = np.ascontiguousarray(keys, dtype=np.uint32).reshape(len(rays), 3)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        self.L.orc_trace_closest_after.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_uint32, C.c_uint32, C.c_void_p]
        self.L.orc_trace_closest_after.restype = None
        t = keys[:, 0].copy().view(np.float32)
        for i in range(len(rays)):
            self.L.orc_trace_closest_after(self.h, C.c_void_p(rays.ctypes.data + i * RAY_DTYPE.itemsize), C.c_float(float(t[i])), int(keys[i, 1]), int(keys[i, 2]),
                                           C.c_void_p(hits.ctypes.data + i * HIT_DTYPE.itemsize))
        return hits

    def generate_primary(self, sys, launch_width, launch_height, iteration):
        rays = np.zeros(launch_width * launch_height, dtype=RAY_DTYPE)
        self.L.orc_generate_primary(self.h, C.byref(sys), launch_width, launch_height, iteration, _ptr(rays))
        return rays

    def render(self, sys, miss, launch_width, launch_height, local_copy=False, iter_first=0, iter_count=1, row_step=1, row_offset=0,
               threads=0, buffer=None, stats=None):
        if buffer is None:
            n = launch_width * launch_height if local_copy else sys.resolution.x * sys.resolution.y
            buffer = np.zeros((n, 4), dtype=np.float32)
        self.L.orc_render(self.h, C.byref(sys), miss, launch_width, launch_height, 1 if local_copy else 0, iter_first, iter_count,
                          row_step, row_offset, threads, _ptr(buffer), C.byref(stats) if stats is not None else None)
        return buffer

    def path_radiance(self, sys, miss, launch_width, launch_xy, iteration, stats=None):
        xy = np.ascontiguousarray(launch_xy, dtype=np.uint32).reshape(-1, 2)
        out = np.zeros((len(xy), 3), dtype=np.float32)
        self.L.orc_path_radiance(self.h, C.byref(sys), miss, launch_width, _ptr(xy), len(xy), iteration, _ptr(out),
                                 C.byref(stats) if stats is not None else None)
        return out


class Reference:
    """The reference's own device programs, host-compiled (oracle/_ref/libref.so), driven one launch index at a time.
    `scene` must be a committed Scene(variant="libm"): it serves optixTrace and keeps the arrays the programs read."""

    def __init__(self, scene, miss):
        if scene.variant != "libm":
            raise ValueError("the reference driver links the libm oracle")
        path = build_reference()
        if path is None:
            raise FileNotFoundError("oracle/_ref/libref.so is not built and /root/reference is absent")
        self.scene = scene
        R = C.CDLL(path)
        R.ref_setup.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_float]
        R.ref_render.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_void_p]
        R.ref_render_rows.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        R.ref_tea4.argtypes = [C.c_uint, C.c_uint]
        R.ref_tea4.restype = C.c_uint
        R.ref_rng.argtypes = [C.POINTER(C.c_uint)]
        R.ref_rng.restype = C.c_float
        self.R = R
        k = scene.keep
        n = len(k["instances"])
        self.attr_ptrs = (C.c_void_p * n)(*[k["geometries"][g][0].ctypes.data for (_, g, _, _) in k["instances"]])
        self.index_ptrs = (C.c_void_p * n)(*[k["geometries"][g][1].ctypes.data for (_, g, _, _) in k["instances"]])
        self.mat = np.array([m for (_, _, m, _) in k["instances"]], dtype=np.int32)
        self.light = np.array([l for (_, _, _, l) in k["instances"]], dtype=np.int32)
        self.xf = np.concatenate([t for (t, _, _, _) in k["instances"]]).astype(np.float32) if n else np.zeros(0, np.float32)
        env = k.get("env")
        lights = k.get("lights")
        R.ref_setup(scene.h, miss, n, self.attr_ptrs, self.index_ptrs, _ptr(self.mat), _ptr(self.light), _ptr(self.xf),
                    _ptr(k["camera"]), _ptr(lights) if lights is not None and len(lights) else None, _ptr(k["materials"]),
                    _ptr(env[0]) if env else None, env[0].shape[1] if env else 0, env[0].shape[0] if env else 0,
                    _ptr(env[1]) if env else None, _ptr(env[2]) if env else None, C.c_float(env[3] if env else 1.0))

    def render(self, sys, launch_width, launch_height, local_copy=False, iter_first=0, iter_count=1, row_step=1, row_offset=0, buffer=None):
        if buffer is None:
            n = launch_width * launch_height if local_copy else sys.resolution.x * sys.resolution.y
            buffer = np.zeros((n, 4), dtype=np.float32)
        self.R.ref_render_rows(C.byref(sys), launch_width, launch_height, 1 if local_copy else 0, iter_first, iter_count, row_step, row_offset,
                               _ptr(buffer))
        return buffer

    def tea4(self, a, b):
        return self.R.ref_tea4(a & 0xffffffff, b & 0xffffffff)

    def rng_sequence(self, seed, n):
        s = C.c_uint(seed)
        return [self.R.ref_rng(C.byref(s)) for _ in range(n)], s.value


def tea4(v0, v1):
    return lib().orc_tea4(v0 & 0xffffffff, v1 & 0xffffffff)


def rng_sequence(seed, n):
    s = C.c_uint32(seed)
    return [lib().orc_rng(C.byref(s)) for _ in range(n)], s.value


def composite(args, tile, out):
    lib().orc_composite(C.byref(args), _ptr(tile), _ptr(out))


def tonemap(params, rgba):
    rgba = np.ascontiguousarray(rgba, dtype=np.float32).reshape(-1, 4)
    out = np.zeros((len(rgba), 3), dtype=np.uint8)
    lib().orc_tonemap(C.byref(params), _ptr(rgba), _ptr(out), len(rgba))
    return out


class WideScene(C.Structure):
    _fields_ = [("tlasNodes", C.c_void_p), ("tlasLeaves", C.c_void_p), ("worldToObject", C.c_void_p), ("instGas", C.c_void_p),
                ("gasNodes", C.c_void_p), ("gasTris", C.c_void_p), ("numInstances", C.c_uint32)]


def wide_trace(export, rays, any_hit=False, variant="pinned", levels=False):
    """Scalar traversal of the PRODUCT's exported wide BVH (core.Context.scene_export) in the product's own order of operations
    (oracle/wide_bvh.inc): returns (hits, (nodes, tris, instances)) -- the work counters the GPU's counting kernels must equal.
    levels=True appends a fourth counter: the nodes that belong to the instance level."""
    L = lib(variant)
    handles = sorted(export["gas"])
    slot = {g: k for k, g in enumerate(handles)}
    inst_gas = np.ascontiguousarray([slot[int(g)] for g in export["instance_gas"]], dtype=np.uint32)
    keep = [np.ascontiguousarray(export["gas"][g][0]) for g in handles], [np.ascontiguousarray(export["gas"][g][1]) for g in handles]
    node_ptrs = (C.c_void_p * max(len(handles), 1))(*[a.ctypes.data for a in keep[0]])
    tri_ptrs = (C.c_void_p * max(len(handles), 1))(*[a.ctypes.data for a in keep[1]])
    tn, tl, w2o = (np.ascontiguousarray(export[k]) for k in ("tlas_nodes", "tlas_leaves", "world_to_object"))
    ws = WideScene(tn.ctypes.data, tl.ctypes.data, w2o.ctypes.data, inst_gas.ctypes.data, C.cast(node_ptrs, C.c_void_p), C.cast(tri_ptrs, C.c_void_p),
                   len(export["instance_gas"]))
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    hits = np.zeros(len(rays), dtype=HIT_DTYPE)
    counts = (C.c_uint64 * 4)()
    fn = L.orc_wide_trace_levels if levels else L.orc_wide_trace
    fn.argtypes = [C.POINTER(WideScene), C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.POINTER(C.c_uint64)]
    fn.restype = None
    fn(C.byref(ws), _ptr(rays), len(rays), 1 if any_hit else 0, _ptr(hits), counts)
    return hits, tuple(int(counts[k]) for k in range(4 if levels else 3))


def online_cores():
    return lib().orc_online_cores()
