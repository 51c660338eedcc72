#!/usr/bin/env python
"""TEST INFRASTRUCTURE (it drives the test oracle and the host build of the traversal source; nothing in the product imports it).

SIMD-aware traversal cost of a scene's acceleration structure WITHOUT a GPU.

tests/tools/bvh_quality.py counts work per RAY.  A warp pays per ITERATION: the node visit once if any lane visits a node, the
triangle loop for as long as the lane with the most triangles, the instance entry once if any lane enters one.  This tool runs
the product's traversal source compiled for the host (tests/native/trace_host.cpp) as one persistent warp in lock step --
32 lanes, refill when 12 are idle, rays in the order the GPU's queues hold them -- over the structure the host-only twin of the
builder makes, and prints, per ray set (primary rays in 8x4 pixel tiles; bounce rays from the primary hit points grouped by
direction octant inside runs of 256 like block_append2_sorted does; the same rays as shadow rays):
  iterations, node passes, triangle passes (sum of per-iteration maxima), instance passes, mean live lanes per iteration,
  and cost = cN * node passes + cT * triangle passes + cI * instance passes + c0 * iterations + cF * refills, per ray.
The constants are static instruction counts of the kernel's phases relative to a node visit (profiles/sass_mix_r1.txt:
node visit ~ 230 instructions, one triangle test ~ 100, instance entry ~ 170, loop + pop ~ 40, ray fetch + begin + store ~ 140).  It is a MODEL -- no memory
system, no scheduler -- meant for A/B comparisons between builder settings (RTC_HOST_LEAF_MAX, RTC_TLAS_LEAF,
RTC_HOST_COLLAPSE, RTC_INSTANCE_BOUNDS ...); profiles/bvh_quality_r2.md compares it with the A/Bs round 2 measured on a B200.

  python tests/tools/simd_cost.py [--config c1|c2|c4|textures] [--width 240 --height 136] [--instances 10000]
                            [--threshold 12] [--leaf-threshold 0] [--define RTC_ONE_TRI_PER_STEP=1]
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))

import helpers as H                                # noqa: E402
from oracle import orc                             # noqa: E402
from tweeker_raytracer_b200 import core            # noqa: E402
import bvh_quality                                 # noqa: E402

C_NODE, C_TRI, C_INST, C_ITER, C_FETCH = 1.0, 0.45, 0.75, 0.17, 0.6


def tile_order(w, h):
    """launch_xy of kernels_shade.cu: consecutive path ids cover 8x4 pixel tiles."""
    idx = np.arange(w * h)
    tile, within = idx >> 5, idx & 31
    tpr = w >> 3
    ty, tx = tile // tpr, tile % tpr
    return ((ty << 2) + (within >> 3)) * w + (tx << 3) + (within & 7)


def octant_runs(rays, run=256):
    """Entries of a CTA's run grouped by the direction octant of their ray (block_append2_sorted)."""
    octant = ((rays["dx"] < 0) * 4 + (rays["dy"] < 0) * 2 + (rays["dz"] < 0)).astype(np.int64)
    key = (np.arange(len(rays)) // run) * 8 + octant
    return rays[np.argsort(key, kind="stable")]


def summarise(c):
    rays = max(c["rays"], 1)
    cost = C_NODE * c["node_passes"] + C_TRI * c["tri_passes_max"] + C_INST * c["inst_passes"] + C_ITER * c["iterations"] + C_FETCH * c["refills"]
    return {"rays": c["rays"], "iterations_per_ray": round(c["iterations"] / rays, 4), "live_lanes": round(c["lane_steps"] / max(c["iterations"], 1), 2),
            "node_passes_per_ray": round(c["node_passes"] / rays, 4), "tri_passes_per_ray": round(c["tri_passes_max"] / rays, 4),
            "inst_passes_per_ray": round(c["inst_passes"] / rays, 4), "refills_per_ray": round(c["refills"] / rays, 4), "nodes": round(c["nodes"] / rays, 3), "tris": round(c["tris"] / rays, 3),
            "instances": round(c["instances"] / rays, 3), "warp_cost_per_ray": round(cost / rays, 4)}


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config", default="c2", choices=sorted(bvh_quality.CONFIGS))
    ap.add_argument("--width", type=int, default=240)
    ap.add_argument("--height", type=int, default=136)
    ap.add_argument("--instances", type=int, default=10000)
    ap.add_argument("--threshold", type=int, default=12, help="idle lanes that trigger a refill (RTC_FETCH_THRESHOLD)")
    ap.add_argument("--leaf-threshold", type=int, default=0, help="> 0: the gated leaf phase of trace_stream (RTC_LEAF_THRESHOLD)")
    ap.add_argument("--define", action="append", default=[], help="compile-time switch of csrc/trace.cuh for the host build, e.g. RTC_ONE_TRI_PER_STEP=1")
    args = ap.parse_args()
    assert args.width % 8 == 0 and args.height % 4 == 0
    with tempfile.TemporaryDirectory() as tmp:
        app = bvh_quality.load(args.config, args.width, args.height, args.instances, tmp)
        try:
            geos = [app.geometry(g) for g in range(app.info.numGeometries)]
            insts = [app.instance(i)[:2] for i in range(app.info.numInstances)]
            export, info = core.host_scene_export(geos, insts)
            ref = H.oracle_scene(app)
            w, h = app.resolution
            primary = ref.generate_primary(H.oracle_sys(app), w, h, 0)[tile_order(w, h)]
            hits = ref.trace_closest(primary)
            bounce = octant_runs(bvh_quality.bounce_rays(primary, hits, 7))
        finally:
            app.close()
    result = {"config": args.config, "resolution": [args.width, args.height], "nodes_total": sum(info["gas_nodes"]) + info["tlas_nodes"]}
    defs = tuple(args.define)
    result["primary"] = summarise(H.product_simd_cost(export, primary, False, args.threshold, defs, args.leaf_threshold))
    result["bounce"] = summarise(H.product_simd_cost(export, bounce, False, args.threshold, defs, args.leaf_threshold))
    result["shadow"] = summarise(H.product_simd_cost(export, bounce, True, args.threshold, defs, args.leaf_threshold))
    print(json.dumps(result))


if __name__ == "__main__":
    main()
