#!/usr/bin/env python
"""Generates tests/golden/wide_bvh_small.npz ON A GPU BOX: the wide BVH the product builds for
tests/golden/scene_small_wide_bvh.txt (rtc_scene_export), a seeded ray set, and the GPU's own answers -- closest hits, any
hits and the work counters of the counting kernels.  The CPU suite then checks, without a GPU, that the oracle's traversal of
this exported structure (oracle/wide_bvh.inc) reproduces those counters and hits, and that the hits equal those of the oracle's
own binary BVH over the same scene.   usage (under gpurun): python tests/golden/make_golden_wide_bvh.py gpurun_out/wide_bvh_small.npz"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H  # noqa: E402
from tweeker_raytracer_b200 import host  # noqa: E402

SCENE = os.path.join(HERE, "scene_small_wide_bvh.txt")


def rays_for_fixture():
    a = H.random_rays(6000, seed=20261018, lo=(-2.5, 0.05, -2.5), hi=(2.5, 2.5, 2.5))
    b = H.random_rays(2000, seed=7, lo=(-2.5, 0.05, -2.5), hi=(2.5, 2.5, 2.5), tmax=1.2)
    return np.concatenate([a, b])


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "wide_bvh_small.npz")
    tmp = tempfile.mkdtemp()
    app = host.App(H.write_system(tmp, "rtigo3_cornell_box", resolution="32 32", samplesSqrt=1), SCENE)
    ctx = app.context(0)
    top = app.system_data(0).topObject
    export = ctx.scene_export(top)
    rays = rays_for_fixture()
    hits = ctx.trace_closest_host(top, rays)
    occluded = ctx.trace_any_host(top, rays)
    d_rays = ctx.to_device(rays)
    cc = ctx.trace_count(top, d_rays, len(rays), any_hit=False)
    ca = ctx.trace_count(top, d_rays, len(rays), any_hit=True)
    ctx.free(d_rays)
    data = {"tlas_nodes": export["tlas_nodes"], "tlas_leaves": export["tlas_leaves"], "world_to_object": export["world_to_object"],
            "instance_gas": export["instance_gas"], "rays": rays, "gpu_hits": hits, "gpu_occluded": occluded.astype(np.uint8),
            "gpu_counts_closest": np.array([cc.nodes, cc.tris, cc.instances, cc.rays], dtype=np.uint64),
            "gpu_counts_any": np.array([ca.nodes, ca.tris, ca.instances, ca.rays], dtype=np.uint64)}
    for g, (nodes, tris) in export["gas"].items():
        data["gas%d_nodes" % g] = nodes
        data["gas%d_tris" % g] = tris
    np.savez_compressed(out, **data)
    app.close()
    print("wrote", out, {k: getattr(v, "shape", None) for k, v in data.items()})


if __name__ == "__main__":
    main()
