"""The PRODUCT's traversal source on the CPU: csrc/trace.cuh -- node test, watertight triangle test, instance entry, the stack
machine of Traversal::step, the SKIP variant of the ordered any-hit processing -- compiled unchanged by g++
(tests/native/trace_host.cpp gives the CUDA intrinsics their IEEE meaning; one host thread plays one lane) and held against

  * a B200: on the structure a B200 exported it finds the GPU's hits and counts the GPU's nodes / triangles / instance
    entries (tests/golden/wide_bvh_small.npz), closest hit and any hit;
  * the scalar oracle: hits bit for bit equal to the oracle's own binary BVH and to brute force over every triangle, counters
    equal to oracle/wide_bvh.inc -- which makes that restatement a checked model of the kernels on a machine without a GPU;
  * the ordered any-hit enumeration: repeated SKIP queries list a ray's candidates in the canonical order
    (t, instance, primitive) exactly as orc_trace_closest_after does;
  * today's builder: the same on structures built by the host-only twin of the builder for every leaf size and both collapses,
    with no traversal-stack overflow.

What this does not execute is the warp-level driver (trace_stream: ballots, the ray cursor) and rcp.approx (the host build
uses the IEEE reciprocal, like the GPU's counting kernels)."""
import os

import numpy as np
import pytest

import helpers as H
from oracle import orc
from tweeker_raytracer_b200 import core, host

GOLD = os.path.join(H.ROOT, "tests", "golden")


def fixture_export():
    z = np.load(os.path.join(GOLD, "wide_bvh_small.npz"))
    gas = {}
    for key in z.files:
        if key.startswith("gas") and key.endswith("_nodes"):
            g = int(key[3:-6])
            gas[g] = (z[key], z["gas%d_tris" % g])
    return z, {"tlas_nodes": z["tlas_nodes"], "tlas_leaves": z["tlas_leaves"], "world_to_object": z["world_to_object"],
               "instance_gas": z["instance_gas"], "gas": gas}


def small_scene(tmp_path):
    app = host.App(H.write_system(tmp_path, "rtigo3_cornell_box", resolution="32 32", samplesSqrt=1),
                   os.path.join(GOLD, "scene_small_wide_bvh.txt"), host_only=True)
    geos = [app.geometry(g) for g in range(app.info.numGeometries)]
    insts = [app.instance(i)[:2] for i in range(app.info.numInstances)]
    return app, geos, insts


def test_host_build_of_the_traversal_source_equals_the_b200(built):
    z, export = fixture_export()
    hits, counts, overflows = H.product_trace(export, z["rays"])
    assert H.hits_equal(hits, z["gpu_hits"])
    assert counts == tuple(int(v) for v in z["gpu_counts_closest"][:3])
    occl, counts_any, _ = H.product_trace(export, z["rays"], any_hit=True)
    assert np.array_equal(occl["inst"] != 0xffffffff, z["gpu_occluded"].astype(bool))
    assert counts_any == tuple(int(v) for v in z["gpu_counts_any"][:3])
    assert overflows == 0


def test_traversal_source_equals_the_oracle_and_its_restatement(built, tmp_path):
    z, export = fixture_export()
    app, _, _ = small_scene(tmp_path)
    ref = H.oracle_scene(app)
    rays = np.concatenate([z["rays"], H.random_rays(3000, 11)])
    hits, counts, _ = H.product_trace(export, rays)
    assert H.hits_equal(hits, ref.trace_closest(rays))
    assert H.hits_equal(hits[:400], ref.trace_closest(rays[:400], brute_force=True))
    model_hits, model_counts = orc.wide_trace(export, rays)
    assert H.hits_equal(hits, model_hits) and counts == model_counts
    occl, counts_any, _ = H.product_trace(export, rays, any_hit=True)
    model_occl, model_counts_any = orc.wide_trace(export, rays, any_hit=True)
    assert counts_any == model_counts_any
    assert np.array_equal(occl["inst"], model_occl["inst"]) and np.array_equal(occl["prim"], model_occl["prim"])
    assert np.array_equal(occl["inst"] != 0xffffffff, ref.trace_any(rays).astype(bool))
    app.close()


def test_skip_variant_enumerates_candidates_in_canonical_order(built, tmp_path):
    z, export = fixture_export()
    app, _, _ = small_scene(tmp_path)
    ref = H.oracle_scene(app)
    rays = z["rays"][:600].copy()
    prev = ref.trace_closest(rays)
    assert H.hits_equal(H.product_trace(export, rays)[0], prev)
    listed = 0
    for _ in range(4):                       # candidate 2, 3, 4, 5 of every ray that still has one
        live = prev["inst"] != 0xffffffff
        if not live.any():
            break
        rays, prev = rays[live], prev[live]
        keys = np.stack([prev["t"].view(np.uint32), prev["inst"], prev["prim"]], axis=1)
        want = ref.trace_closest_after(rays, keys)
        got, _, _ = H.product_trace(export, rays, skip=keys)
        assert H.hits_equal(got, want)
        found = got["inst"] != 0xffffffff
        # strictly after the key in (t, instance, primitive)
        later = (got["t"] > prev["t"]) | ((got["t"] == prev["t"]) & ((got["inst"] > prev["inst"]) | ((got["inst"] == prev["inst"]) & (got["prim"] > prev["prim"]))))
        assert later[found].all()
        listed += int(found.sum())
        prev = got
    assert listed > 200
    app.close()


@pytest.mark.parametrize("leaf_max", [1, 2, 3])
@pytest.mark.parametrize("collapse", ["optimal", "greedy"])
def test_traversal_source_on_todays_builder(built, tmp_path, monkeypatch, leaf_max, collapse):
    monkeypatch.setenv("RTC_HOST_LEAF_MAX", str(leaf_max))
    monkeypatch.setenv("RTC_HOST_COLLAPSE", collapse)
    monkeypatch.setenv("RTC_TLAS_COLLAPSE", collapse)
    z, _ = fixture_export()
    app, geos, insts = small_scene(tmp_path)
    export, _ = core.host_scene_export(geos, insts)
    ref = H.oracle_scene(app)
    rays = z["rays"]
    hits, counts, overflows = H.product_trace(export, rays)
    assert overflows == 0
    assert H.hits_equal(hits, ref.trace_closest(rays)) and H.hits_equal(hits, z["gpu_hits"])
    assert counts == orc.wide_trace(export, rays)[1]
    occl, counts_any, _ = H.product_trace(export, rays, any_hit=True)
    assert np.array_equal(occl["inst"] != 0xffffffff, z["gpu_occluded"].astype(bool))
    assert counts_any == orc.wide_trace(export, rays, any_hit=True)[1]
    app.close()


def test_traversal_source_on_the_cornell_box_and_an_instanced_lattice(built, tmp_path):
    """Larger structures of today's builder: the Cornell box (32 K-triangle sphere) and 343 rotated instances of a torus."""
    import sys
    sys.path.insert(0, os.path.join(H.ROOT, "tools"))
    import make_instances_scene
    lattice = os.path.join(str(tmp_path), "scene_lattice.txt")
    make_instances_scene.write_scene(lattice, count=343, tess=(24, 12))
    for scene, name, box in ((H.scene_path("rtigo3_cornell_box"), "rtigo3_cornell_box", None), (lattice, "rtigo3_instances", 12.0)):
        app = host.App(H.write_system(tmp_path, name, resolution="48 32", samplesSqrt=1), scene, host_only=True)
        geos = [app.geometry(g) for g in range(app.info.numGeometries)]
        insts = [app.instance(i)[:2] for i in range(app.info.numInstances)]
        export, _ = core.host_scene_export(geos, insts)
        ref = H.oracle_scene(app)
        w, h = app.resolution
        primary = ref.generate_primary(H.oracle_sys(app), w, h, 0)
        rand = H.random_rays(2500, 5) if box is None else H.random_rays(2500, 5, lo=(-box, 0.0, -box), hi=(box, box, box))
        rays = np.concatenate([primary[primary["tmax"] > 0], rand])
        hits, counts, overflows = H.product_trace(export, rays)
        assert overflows == 0
        assert H.hits_equal(hits, ref.trace_closest(rays))
        assert counts == orc.wide_trace(export, rays)[1]
        assert 0.05 < float((hits["inst"] != 0xffffffff).mean())
        app.close()


@pytest.mark.parametrize("seed", list(range(16)))
def test_random_scenes_of_the_gpu_fuzz_test(built, tmp_path, seed):
    """The scenes of tests/test_gpu_fuzz.py (same generator, same seeds: random instance transforms with non-uniform scale and
    rotation, every model kind) through the host-only builder and the host build of the traversal source: closest hits and
    occlusion of the same 30 000 random rays equal the oracle's, and the counters equal the restatement's."""
    from test_gpu_fuzz import random_scene
    rng = np.random.default_rng(1000 + seed)
    scene = os.path.join(str(tmp_path), "scene_fuzz.txt")
    random_scene(scene, rng)
    app = host.App(H.write_system(tmp_path, "rtigo3_geometry", resolution="32 32", samplesSqrt=1), scene, host_only=True)
    geos = [app.geometry(g) for g in range(app.info.numGeometries)]
    insts = [app.instance(i)[:2] for i in range(app.info.numInstances)]
    export, _ = core.host_scene_export(geos, insts)
    ref = H.oracle_scene(app)
    rays = H.random_rays(30000, seed=seed, lo=(-5, 0.05, -5), hi=(5, 4, 5))
    hits, counts, overflows = H.product_trace(export, rays)
    assert overflows == 0
    assert H.hits_equal(hits, ref.trace_closest(rays))
    assert counts == orc.wide_trace(export, rays)[1]
    occl, _, _ = H.product_trace(export, rays, any_hit=True)
    assert np.array_equal(occl["inst"] != 0xffffffff, ref.trace_any(rays).astype(bool))
    app.close()


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_triangle_soups_with_degenerate_input(built, seed):
    """Unstructured soups straight through the C ABI of the host-only builder: triangles of wildly different sizes, exact
    duplicates, zero-area triangles (repeated vertex, collinear), coordinates far from the origin; two instances (one mirrored
    and non-uniformly scaled).  The reference for the hits is brute force over every triangle."""
    rng = np.random.default_rng(seed)
    n = 1500
    centre = rng.uniform(-3, 3, size=(n, 1, 3)) + (1000.0 if seed == 4 else 0.0)
    size = np.exp(rng.uniform(np.log(1e-3), np.log(2.0), size=(n, 1, 1)))
    tris = (centre + size * rng.normal(size=(n, 3, 3))).astype(np.float32)
    tris[10:20] = tris[0:10]                       # exact duplicates: ties -> smaller primitive id
    tris[20:30, 1] = tris[20:30, 0]                # repeated vertex
    tris[30:40, 2] = (tris[30:40, 0] + tris[30:40, 1]) * np.float32(0.5)      # collinear
    verts = tris.reshape(-1, 3)
    idx = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float32)
    warped = np.array([-0.7, 0.2, 0, 1.5, 0.1, 1.3, 0.3, -0.5, 0, -0.4, 0.6, 2.0], dtype=np.float32)
    export, info = core.host_scene_export([(verts, idx)], [(ident, 0), (warped, 0)])
    attrs = np.zeros(len(verts), dtype=orc.ATTR_DTYPE)
    attrs["vertex"] = verts
    ref = orc.Scene()
    ref.add_geometry(attrs, idx)
    ref.add_instance(ident, 0, 0)
    ref.add_instance(warped, 0, 0)
    ref.commit()
    off = 1000.0 if seed == 4 else 0.0
    rays = H.random_rays(4000, seed=seed, lo=(-4 + off, -4 + off, -4 + off), hi=(4 + off, 4 + off, 4 + off))
    hits, counts, overflows = H.product_trace(export, rays)
    assert overflows == 0
    assert H.hits_equal(hits, ref.trace_closest(rays, brute_force=True))
    assert H.hits_equal(hits, ref.trace_closest(rays))
    assert counts == orc.wide_trace(export, rays)[1]
    assert float((hits["inst"] != 0xffffffff).mean()) > 0.2
    ref.close()


def test_lock_step_warp_model_is_consistent_with_the_per_ray_counters(built):
    """tests/tools/simd_cost.py's instrument (th_simd_cost): one persistent warp in lock step.  Its summed lane work must be the work
    the per-ray run counts, whatever the refill threshold; its per-iteration maxima are bounded by both."""
    z, export = fixture_export()
    rays = z["rays"]
    _, counts, _ = H.product_trace(export, rays)
    for threshold in (1, 12, 32):
        c = H.product_simd_cost(export, rays, fetch_threshold=threshold)
        assert (c["nodes"], c["tris"], c["instances"]) == counts and c["rays"] == len(rays)
        assert c["node_passes"] <= c["iterations"] and c["inst_passes"] <= c["iterations"]
        assert c["nodes"] <= 32 * c["node_passes"] and c["node_passes"] <= c["nodes"]
        assert c["tris"] <= 32 * c["tri_passes_max"] and c["tri_passes_max"] <= c["tris"]
        assert c["lane_steps"] <= 32 * c["iterations"]
    eager, lazy = H.product_simd_cost(export, rays, fetch_threshold=1), H.product_simd_cost(export, rays, fetch_threshold=32)
    assert eager["iterations"] < lazy["iterations"]          # refilling early keeps the lanes busy: fewer warp iterations
    occl = H.product_simd_cost(export, rays, any_hit=True)
    assert (occl["nodes"], occl["tris"], occl["instances"]) == H.product_trace(export, rays, any_hit=True)[1]


@pytest.mark.parametrize("cap", [1, 2])
def test_capped_schedules_change_no_hit_and_no_counter(built, tmp_path, cap):
    """The capped schedules of the triangle tests (Traversal<..., TRICAP = 1 | 2>, RTC_SCHEDULE_ONE_TRI / RTC_SCHEDULE_TWO_TRI; the
    host build selects them with -DRTC_ONE_TRI_PER_STEP=<cap>, the default of the template parameter): the same tests in the same
    order per ray -- identical hits, identical work counters, closest hit, any hit and the SKIP enumeration -- spread over more
    iterations with at most <cap> triangles each.  The GPU twin of this test is tests/test_gpu_trace_schedule.py."""
    defs = ("RTC_ONE_TRI_PER_STEP=%d" % cap,)
    z, export = fixture_export()
    rays = np.concatenate([z["rays"], H.random_rays(3000, 3)])
    for any_hit in (False, True):
        a_hits, a_counts, _ = H.product_trace(export, rays, any_hit=any_hit)
        b_hits, b_counts, overflows = H.product_trace(export, rays, any_hit=any_hit, defs=defs)
        assert H.hits_equal(a_hits, b_hits) and a_counts == b_counts and overflows == 0
    first = H.product_trace(export, rays[:500])[0]
    live = first["inst"] != 0xffffffff
    keys = np.stack([first["t"].view(np.uint32), first["inst"], first["prim"]], axis=1)[live]
    assert H.hits_equal(H.product_trace(export, rays[:500][live], skip=keys)[0], H.product_trace(export, rays[:500][live], skip=keys, defs=defs)[0])
    base, variant = H.product_simd_cost(export, rays), H.product_simd_cost(export, rays, defs=defs)
    assert (variant["nodes"], variant["tris"], variant["instances"]) == (base["nodes"], base["tris"], base["instances"])
    assert variant["tri_passes_max"] < (0.6, 0.8)[cap - 1] * base["tri_passes_max"] and variant["iterations"] > base["iterations"]


@pytest.mark.parametrize("seed", [0, 5, 11])
@pytest.mark.parametrize("cap", [1, 2])
def test_capped_schedules_on_random_scenes(built, tmp_path, seed, cap):
    """The capped schedules on scenes of the GPU fuzz test (rotated, non-uniformly scaled instances of every model kind): hits,
    occlusion and work counters equal those of the uncapped schedule and of the oracle."""
    from test_gpu_fuzz import random_scene
    defs = ("RTC_ONE_TRI_PER_STEP=%d" % cap,)
    rng = np.random.default_rng(1000 + seed)
    scene = os.path.join(str(tmp_path), "scene_fuzz.txt")
    random_scene(scene, rng)
    app = host.App(H.write_system(tmp_path, "rtigo3_geometry", resolution="32 32", samplesSqrt=1), scene, host_only=True)
    geos = [app.geometry(g) for g in range(app.info.numGeometries)]
    insts = [app.instance(i)[:2] for i in range(app.info.numInstances)]
    export, _ = core.host_scene_export(geos, insts)
    ref = H.oracle_scene(app)
    rays = H.random_rays(20000, seed=seed, lo=(-5, 0.05, -5), hi=(5, 4, 5))
    hits, counts, overflows = H.product_trace(export, rays, defs=defs)
    assert overflows == 0 and H.hits_equal(hits, ref.trace_closest(rays))
    assert counts == H.product_trace(export, rays)[1] == orc.wide_trace(export, rays)[1]
    occl, counts_any, _ = H.product_trace(export, rays, any_hit=True, defs=defs)
    assert np.array_equal(occl["inst"] != 0xffffffff, ref.trace_any(rays).astype(bool))
    assert counts_any == orc.wide_trace(export, rays, any_hit=True)[1]
    app.close()
