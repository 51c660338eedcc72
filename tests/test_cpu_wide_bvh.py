"""CPU check of the oracle's traversal of the PRODUCT's wide BVH (oracle/wide_bvh.inc) against a committed fixture:
tests/golden/wide_bvh_small.npz holds the acceleration structure a B200 built for tests/golden/scene_small_wide_bvh.txt
(rtc_scene_export), a seeded ray set and the GPU's own hits and work counters (tests/golden/make_golden_wide_bvh.py).

  * the oracle walking the exported structure counts exactly the nodes / triangles / instance entries the GPU counted;
  * it finds exactly the GPU's hits;
  * those hits equal the hits of the oracle's OWN binary BVH built from the scene file: the closest hit does not depend on the
    acceleration structure, which is what lets a scalar binary-BVH intersector stand in for optixTrace in the first place."""
import os

import numpy as np

import helpers as H
from oracle import orc
from tweeker_raytracer_b200 import host

GOLD = os.path.join(H.ROOT, "tests", "golden")


def load_fixture():
    z = np.load(os.path.join(GOLD, "wide_bvh_small.npz"))
    gas = {}
    for key in z.files:
        if key.startswith("gas") and key.endswith("_nodes"):
            g = int(key[3:-6])
            gas[g] = (z[key], z["gas%d_tris" % g])
    export = {"tlas_nodes": z["tlas_nodes"], "tlas_leaves": z["tlas_leaves"], "world_to_object": z["world_to_object"],
              "instance_gas": z["instance_gas"], "gas": gas}
    return z, export


def test_oracle_counts_and_hits_equal_the_gpu_on_the_exported_wide_bvh(built):
    z, export = load_fixture()
    rays = z["rays"]
    hits, counts = orc.wide_trace(export, rays)
    assert counts == tuple(int(v) for v in z["gpu_counts_closest"][:3])
    assert H.hits_equal(hits, z["gpu_hits"])
    occl, counts_any = orc.wide_trace(export, rays, any_hit=True)
    assert counts_any == tuple(int(v) for v in z["gpu_counts_any"][:3])
    assert np.array_equal(occl["inst"] != 0xffffffff, z["gpu_occluded"].astype(bool))
    assert 0.2 < float((hits["inst"] != 0xffffffff).mean()) < 1.0 and counts[0] > len(rays)       # the fixture is not trivial


def test_wide_bvh_hits_equal_the_binary_bvh_of_the_scene_file(built, tmp_path):
    z, export = load_fixture()
    app = host.App(H.write_system(tmp_path, "rtigo3_cornell_box", resolution="32 32", samplesSqrt=1),
                   os.path.join(GOLD, "scene_small_wide_bvh.txt"), host_only=True)
    ref = H.oracle_scene(app)
    rays = z["rays"]
    hits, _ = orc.wide_trace(export, rays)
    assert H.hits_equal(hits, ref.trace_closest(rays))
    assert H.hits_equal(hits[:400], ref.trace_closest(rays[:400], brute_force=True))      # ground truth: every triangle of every instance
    assert np.array_equal(z["gpu_occluded"].astype(bool), ref.trace_any(rays).astype(bool))
    for i in range(app.info.numInstances):                                                  # the exported world-to-object matrices are the oracle's
        assert export["world_to_object"][i].tobytes() == ref.inverse(i).tobytes()
    app.close()
