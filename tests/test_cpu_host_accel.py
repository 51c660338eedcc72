"""CPU tests of the acceleration-structure builder through its host-only twin (rtc_host_gas_build / rtc_host_ias_build,
csrc/accel_host.cpp): the same code rtc_gas_build(RTC_BUILD_HOST_SAH) and rtc_ias_build run before their uploads, fed from
host arrays -- no CUDA call, no context.

  * pinned against a B200: with the leaf size of the committed fixture's vintage (<= 3 triangles; today <= 2) the host-only
    build reproduces BYTE FOR BYTE the structure a B200 exported (tests/golden/wide_bvh_small.npz,
    tests/golden/make_golden_wide_bvh.py) -- nodes, leaf-ordered triangles, instance-level leaves, world->object matrices,
    including the instance bounds the device kernel k_instance_bounds computed;
  * every leaf size x both collapses (the greedy default and the opt-in SAH-optimal one): every structural invariant of the node format the kernels
    rely on, and closest hits / occlusion identical to the oracle's own binary BVH and to brute force over every triangle;
  * the optimal collapse never costs more summed node area than the greedy one, and needs fewer nodes;
  * edge cases: empty geometry, one triangle, an instance of an empty mesh, coincident primitives, index out of range.
"""
import os

import numpy as np
import pytest

import helpers as H
from oracle import orc
from tweeker_raytracer_b200 import core, host

GOLD = os.path.join(H.ROOT, "tests", "golden")

NODE_DTYPE = np.dtype([("p", "f4", 3), ("e", "u1", 3), ("imask", "u1"), ("childBase", "u4"), ("triBase", "u4"), ("meta", "u1", 8),
                       ("qlo", "u1", (3, 8)), ("qhi", "u1", (3, 8))])
assert NODE_DTYPE.itemsize == 80


def scene_arrays(tmp_path, scene_file, name="rtigo3_cornell_box"):
    app = host.App(H.write_system(tmp_path, name, resolution="32 32", samplesSqrt=1), scene_file, host_only=True)
    geos = [app.geometry(g) for g in range(app.info.numGeometries)]
    insts = [app.instance(i)[:2] for i in range(app.info.numInstances)]
    return app, geos, insts


def decode(nodes):
    """[n, 80] uint8 -> structured view; qlo / qhi come out as [axis][slot] (x: qlox .. z: qloz, then qhix .. qhiz)."""
    raw = np.ascontiguousarray(nodes).view(np.uint8).reshape(-1, 80)
    out = np.zeros(len(raw), dtype=NODE_DTYPE)
    out["p"] = raw[:, 0:12].copy().view(np.float32).reshape(-1, 3)
    out["e"] = raw[:, 12:15]
    out["imask"] = raw[:, 15]
    out["childBase"] = raw[:, 16:20].copy().view(np.uint32).reshape(-1)
    out["triBase"] = raw[:, 20:24].copy().view(np.uint32).reshape(-1)
    out["meta"] = raw[:, 24:32]
    out["qlo"] = raw[:, 32:56].reshape(-1, 3, 8)
    out["qhi"] = raw[:, 56:80].reshape(-1, 3, 8)
    return out


def check_structure(nodes, num_prims, prim_boxes_in_leaf_order, leaf_max):
    """The invariants trace.cuh relies on: every node reachable exactly once, inner children in one contiguous block addressed by
    the popcount of imask, leaf runs inside the 32-bit primitive mask and disjoint, empty slots with inverted boxes, every
    primitive in exactly one leaf, and every (decoded, conservative) child box containing the primitives below it."""
    n = decode(nodes)
    seen_nodes = np.zeros(len(n), dtype=bool)
    seen_prims = np.zeros(max(num_prims, 1), dtype=np.int32)
    stack = [0]
    # boxes of a node's subtree (decoded, conservative) must contain the primitives below it
    def child_box(node, s):
        step = np.ldexp(np.float32(1.0), node["e"].astype(np.int32) - 127).astype(np.float64)
        lo = node["p"].astype(np.float64) + node["qlo"][:, s].astype(np.float64) * step
        hi = node["p"].astype(np.float64) + node["qhi"][:, s].astype(np.float64) * step
        return lo, hi
    def prims_below(idx, acc):
        node = n[idx]
        inner = 0
        for s in range(8):
            if node["imask"] >> s & 1:
                prims_below(int(node["childBase"]) + inner, acc)
                inner += 1
            elif node["meta"][s]:
                cnt, off = int(node["meta"][s]) >> 5, int(node["meta"][s]) & 31
                acc.extend(range(int(node["triBase"]) + off, int(node["triBase"]) + off + cnt))
    while stack:
        idx = stack.pop()
        assert not seen_nodes[idx], "node referenced twice"
        seen_nodes[idx] = True
        node = n[idx]
        inner = 0
        offsets = []
        for s in range(8):
            is_inner = bool(node["imask"] >> s & 1)
            meta = int(node["meta"][s])
            if is_inner:
                child = int(node["childBase"]) + inner
                inner += 1
                assert child < len(n)
                stack.append(child)
                below = []
                prims_below(child, below)
            elif meta:
                cnt, off = meta >> 5, meta & 31
                assert 1 <= cnt <= leaf_max and off + cnt <= 32
                offsets.append((off, cnt))
                below = list(range(int(node["triBase"]) + off, int(node["triBase"]) + off + cnt))
                for q in below:
                    seen_prims[q] += 1
            else:
                assert all(node["qlo"][k][s] == 255 and node["qhi"][k][s] == 0 for k in range(3)), "empty slot must have an inverted box"
                continue
            lo, hi = child_box(node, s)
            if prim_boxes_in_leaf_order is not None and len(below) <= 64:
                for q in below:
                    assert np.all(lo <= prim_boxes_in_leaf_order[q, 0]) and np.all(hi >= prim_boxes_in_leaf_order[q, 1]), "child box does not contain its primitive"
        offsets.sort()
        for (o0, c0), (o1, _) in zip(offsets, offsets[1:]):
            assert o0 + c0 <= o1, "leaf runs of a node overlap"
    assert seen_nodes.all(), "unreachable node"
    if num_prims:
        assert (seen_prims[:num_prims] == 1).all(), "every primitive must sit in exactly one leaf"
    return int(seen_nodes.sum())


def tri_boxes(tris):
    v = tris.reshape(-1, 3, 4)[:, :, :3].astype(np.float64)
    return np.stack([v.min(axis=1), v.max(axis=1)], axis=1)


def test_host_only_build_reproduces_the_structure_a_b200_exported(built, tmp_path, monkeypatch):
    monkeypatch.setenv("RTC_HOST_LEAF_MAX", "3")          # the leaf size the fixture was exported with (today: 2)
    z = np.load(os.path.join(GOLD, "wide_bvh_small.npz"))
    app, geos, insts = scene_arrays(tmp_path, os.path.join(GOLD, "scene_small_wide_bvh.txt"))
    export, info = core.host_scene_export(geos, insts)
    app.close()
    assert export["tlas_nodes"].tobytes() == z["tlas_nodes"].tobytes()
    assert export["tlas_leaves"].tobytes() == z["tlas_leaves"].tobytes()
    assert export["world_to_object"].tobytes() == z["world_to_object"].tobytes()
    assert np.array_equal(export["instance_gas"], z["instance_gas"])
    for g, (nodes, tris) in export["gas"].items():
        assert nodes.tobytes() == z["gas%d_nodes" % g].tobytes(), g
        assert tris.tobytes() == z["gas%d_tris" % g].tobytes(), g
    hits, counts = orc.wide_trace(export, z["rays"])
    assert counts == tuple(int(v) for v in z["gpu_counts_closest"][:3])
    assert H.hits_equal(hits, z["gpu_hits"])


@pytest.mark.parametrize("leaf_max", [1, 2, 3])
@pytest.mark.parametrize("collapse", ["optimal", "greedy"])
def test_builder_invariants_and_hits(built, tmp_path, monkeypatch, leaf_max, collapse):
    monkeypatch.setenv("RTC_HOST_LEAF_MAX", str(leaf_max))
    monkeypatch.setenv("RTC_HOST_COLLAPSE", collapse)
    monkeypatch.setenv("RTC_TLAS_COLLAPSE", collapse)
    z = np.load(os.path.join(GOLD, "wide_bvh_small.npz"))
    app, geos, insts = scene_arrays(tmp_path, os.path.join(GOLD, "scene_small_wide_bvh.txt"))
    export, info = core.host_scene_export(geos, insts)
    for g, (nodes, tris) in export["gas"].items():
        check_structure(nodes, len(tris), tri_boxes(tris), leaf_max)
        assert sorted(tris[:, 3].view(np.uint32).tolist()) == list(range(len(geos[g][1])))      # primitive ids: a permutation
    check_structure(export["tlas_nodes"], len(export["tlas_leaves"]), None, 1)
    assert sorted(export["tlas_leaves"].tolist()) == list(range(len(insts)))
    ref = H.oracle_scene(app)
    rays = z["rays"]
    hits, _ = orc.wide_trace(export, rays)
    assert H.hits_equal(hits, ref.trace_closest(rays))
    assert H.hits_equal(hits[:300], ref.trace_closest(rays[:300], brute_force=True))
    occl, _ = orc.wide_trace(export, rays, any_hit=True)
    assert np.array_equal(occl["inst"] != 0xffffffff, ref.trace_any(rays).astype(bool))
    assert H.hits_equal(hits, z["gpu_hits"])            # and they are the hits a B200 found on ITS structure
    app.close()


def node_area_sum(nodes):
    """Summed half area of the (decoded) bounds of every wide node: the quantity the optimal collapse minimises."""
    n = decode(nodes)
    total = 0.0
    for node in n:
        used = [s for s in range(8) if (node["imask"] >> s & 1) or node["meta"][s]]
        if not used:
            continue
        step = np.ldexp(np.float32(1.0), node["e"].astype(np.int32) - 127).astype(np.float64)
        lo = (node["qlo"][:, used].astype(np.float64) * step[:, None]).min(axis=1)
        hi = (node["qhi"][:, used].astype(np.float64) * step[:, None]).max(axis=1)
        d = hi - lo
        total += d[0] * d[1] + d[1] * d[2] + d[2] * d[0]
    return total


def test_optimal_collapse_beats_the_greedy_one(built, tmp_path, monkeypatch):
    app, geos, insts = scene_arrays(tmp_path, H.scene_path("rtigo3_cornell_box"))
    app.close()
    result = {}
    for collapse in ("greedy", "optimal"):
        monkeypatch.setenv("RTC_HOST_COLLAPSE", collapse)
        monkeypatch.setenv("RTC_TLAS_COLLAPSE", collapse)
        export, info = core.host_scene_export(geos, insts)
        g = max(export["gas"], key=lambda k: len(export["gas"][k][0]))        # the tessellated sphere
        result[collapse] = (len(export["gas"][g][0]), node_area_sum(export["gas"][g][0]))
    assert result["optimal"][0] < 0.7 * result["greedy"][0]                   # far better filled nodes
    assert result["optimal"][1] <= result["greedy"][1] * 1.001                # quantisation slack aside, never more area


def test_edge_cases(built):
    L = core.lib()
    # empty geometry, and an instance of it: one empty node each, nothing to hit
    empty = (np.zeros((0, 3), dtype=np.float32), np.zeros((0, 3), dtype=np.uint32))
    one = (np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float32), np.array([[0, 1, 2]], dtype=np.uint32))
    # eight coincident triangles (centroid bounds of zero extent: the builder must fall back to a median split)
    same = (np.tile(one[0], (8, 1)), np.arange(24, dtype=np.uint32).reshape(8, 3))
    ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float32)
    shifted = ident.copy(); shifted[3] = 5.0
    export, info = core.host_scene_export([empty, one, same], [(ident, 0), (ident, 1), (shifted, 2)])
    assert info["gas_nodes"][0] == 1 and info["gas_tris"] == [0, 1, 8]
    check_structure(export["gas"][1][0], 1, tri_boxes(export["gas"][1][1]), 3)
    check_structure(export["gas"][2][0], 8, tri_boxes(export["gas"][2][1]), 3)
    rays = np.zeros(3, dtype=orc.RAY_DTYPE)
    rays["oz"], rays["dz"], rays["tmax"] = 1.0, -1.0, 1e27
    rays["ox"] = [0.25, 5.25, -3.0]
    rays["oy"] = 0.25
    hits, counts = orc.wide_trace(export, rays)
    assert hits["inst"].tolist() == [1, 2, 0xffffffff]
    assert hits["prim"].tolist()[:2] == [0, 0]                  # ties between coincident triangles -> the smallest primitive id
    assert np.all(hits["t"][:2] == np.float32(1.0))
    # an index beyond the vertex array is refused, with the message of the device path
    bad = np.array([[0, 1, 7]], dtype=np.uint32)
    import ctypes as C
    h = C.c_void_p()
    rc = L.rtc_host_gas_build(one[0].ctypes.data_as(C.c_void_p), 12, 3, bad.ctypes.data_as(C.c_void_p), 1, C.byref(h))
    assert rc != 0 and b"out of range" in L.rtc_last_error()
    rc = L.rtc_host_gas_build(one[0].ctypes.data_as(C.c_void_p), 10, 3, bad.ctypes.data_as(C.c_void_p), 1, C.byref(h))
    assert rc != 0 and b"stride" in L.rtc_last_error()
