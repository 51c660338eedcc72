#!/usr/bin/env python
"""TEST INFRASTRUCTURE (it drives the test oracle; nothing in the product imports it).

Acceleration-structure quality WITHOUT a GPU: builds the scene's wide BVH through the host-only twin of the builder
(rtc_host_gas_build / rtc_host_ias_build -- byte for byte the arrays a B200 is handed, tests/test_cpu_host_accel.py) and lets
the test oracle traverse it in the kernels' order of operations (oracle/wide_bvh.inc), which reproduces the work counters of
the GPU's counting kernels ray by ray (tests/test_gpu_wide_bvh.py).  Prints wide nodes visited, triangles tested and instances
entered per ray for
  primary : the camera rays of iteration 0 (a subsampled frame),
  bounce  : rays leaving the primary hit points in uniformly random directions (the incoherent continuation / shadow rays),
and the combined cost in node visits (a triangle test costs a warp about 2.7 node visits and an instance entry about 3 in the
lane-owned traversal driver, profiles/sweeps_r2.md).  Every hit is checked against the oracle's own binary BVH.

  python tests/tools/bvh_quality.py [--config c1|c2|c4|textures] [--width 240 --height 135] [--instances 10000] [--any]
Environment knobs of the builder (RTC_HOST_LEAF_MAX, RTC_TLAS_LEAF, RTC_INSTANCE_BOUNDS, RTC_HOST_* ...) apply.
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))

import helpers as H                                # noqa: E402
from oracle import orc                             # noqa: E402
from tweeker_raytracer_b200 import core, host      # noqa: E402

CONFIGS = {"c1": "rtigo3_cornell_box", "c2": "rtigo3_geometry", "c4": "rtigo3_instances", "textures": "rtigo3_textures"}
TRI_COST, INST_COST = 2.7, 3.0


def load(config, width, height, instances, tmp):
    name = CONFIGS[config]
    scene = H.scene_path(name)
    if config == "c4":
        import make_instances_scene
        scene = os.path.join(tmp, "scene_instances.txt")
        make_instances_scene.write_scene(scene, count=instances)
    sysfile = H.write_system(tmp, name, resolution="%d %d" % (width, height), samplesSqrt=1)
    return host.App(sysfile, scene, host_only=True)


def bounce_rays(rays, hits, seed):
    """Uniformly random directions from the primary hit points, pushed off the surface along the new direction."""
    ok = hits["inst"] != 0xffffffff
    r, h = rays[ok], hits[ok]
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(len(r), 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    out = np.zeros(len(r), dtype=orc.RAY_DTYPE)
    for k, (o, dd) in enumerate((("ox", "dx"), ("oy", "dy"), ("oz", "dz"))):
        out[o] = r[o] + h["t"] * r[dd] + np.float32(1e-3) * d[:, k]
        out[dd] = d[:, k]
    out["tmin"] = 1e-4
    out["tmax"] = 1e27
    return out


def measure(app, export, any_hit, seed=7):
    ref = H.oracle_scene(app)
    w, h = app.resolution
    primary = ref.generate_primary(H.oracle_sys(app), w, h, 0)
    primary = primary[primary["tmax"] > 0]
    out = {}
    hits, counts = orc.wide_trace(export, primary, levels=True)
    want = ref.trace_closest(primary)
    if not H.hits_equal(hits, want):
        raise AssertionError("primary hits on the wide BVH differ from the binary BVH of the oracle")
    sets = {"primary": (primary, counts)}
    second = bounce_rays(primary, hits, seed)
    if len(second):
        h2, c2 = orc.wide_trace(export, second, levels=True)
        if not H.hits_equal(h2, ref.trace_closest(second)):
            raise AssertionError("bounce hits on the wide BVH differ from the binary BVH of the oracle")
        sets["bounce"] = (second, c2)
        if any_hit:
            occ, c3 = orc.wide_trace(export, second, any_hit=True, levels=True)
            if not np.array_equal(occ["inst"] != 0xffffffff, ref.trace_any(second).astype(bool)):
                raise AssertionError("occlusion on the wide BVH differs from the binary BVH of the oracle")
            sets["bounce_any"] = (second, c3)
    for name, (rays, c) in sets.items():
        n = max(len(rays), 1)
        per = [c[0] / n, c[1] / n, c[2] / n]
        out[name] = {"rays": len(rays), "nodes": round(per[0], 3), "tris": round(per[1], 3), "instances": round(per[2], 3), "nodes_instance_level": round(c[3] / n, 3),
                     "cost": round(per[0] + TRI_COST * per[1] + INST_COST * per[2], 3)}
    return out


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--width", type=int, default=240)
    ap.add_argument("--height", type=int, default=135)
    ap.add_argument("--instances", type=int, default=10000)
    ap.add_argument("--any", action="store_true", help="also the any-hit (shadow ray) traversal of the bounce rays")
    args = ap.parse_args()
    with tempfile.TemporaryDirectory() as tmp:
        app = load(args.config, args.width, args.height, args.instances, tmp)
        try:
            geos = [app.geometry(g) for g in range(app.info.numGeometries)]
            insts = [app.instance(i)[:2] for i in range(app.info.numInstances)]
            t0 = time.time()
            export, info = core.host_scene_export(geos, insts)
            info["build_s"] = round(time.time() - t0, 3)
            result = {"config": args.config, "resolution": [args.width, args.height], "instances": len(insts), "build": info}
            result.update(measure(app, export, args.any))
        finally:
            app.close()
    print(json.dumps(result))


if __name__ == "__main__":
    main()
