"""The work counters behind the roofline figure, pinned: the scalar oracle traverses the IDENTICAL wide BVH (exported with
rtc_scene_export) in the product's own order of operations (oracle/wide_bvh.inc) and must count exactly the wide nodes, triangles
and instance entries the GPU's counting kernels report -- and find exactly the hits of the oracle's own
binary BVH, because the closest hit does not depend on the acceleration structure."""
import numpy as np
import pytest

import helpers as H
from oracle import orc
from tweeker_raytracer_b200 import core, host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["rtigo3_cornell_box", "rtigo3_geometry"])
def scene(request, cuda_device, tmp_path_factory):
    tmp = tmp_path_factory.mktemp("wide")
    app = host.App(H.write_system(tmp, request.param, resolution="96 64", samplesSqrt=2), H.scene_path(request.param))
    yield request.param, app, H.oracle_scene(app)
    app.close()


def ray_sets(name, app, ref):
    w, h = app.resolution
    box = dict(lo=(-1.0, 0.0, -1.0), hi=(1.0, 2.0, 1.0)) if "cornell" in name else dict(lo=(-8, 0.01, -8), hi=(8, 5, 8))
    primary = ref.generate_primary(H.oracle_sys(app), w, h, 1)
    return {"primary": primary[primary["tmax"] > 0], "incoherent": H.random_rays(20000, seed=0xC0FFEE, **box),
            "short": H.random_rays(20000, seed=0xBEEF, tmax=1.5, **box)}


@pytest.mark.parametrize("any_hit", [False, True])
def test_gpu_work_counters_equal_the_oracle_on_the_exported_bvh(scene, any_hit):
    name, app, ref = scene
    ctx = app.context(0)
    top = app.system_data(0).topObject
    export = ctx.scene_export(top)
    info = ctx.scene_info(top)
    assert export["tlas_nodes"].shape == (info.numTlasNodes, 80) and len(export["instance_gas"]) == info.numInstances
    assert sum(len(n) for n, _ in export["gas"].values()) + info.numTlasNodes == info.numNodes
    for key, rays in ray_sets(name, app, ref).items():
        d_rays = ctx.to_device(rays)
        got = ctx.trace_count(top, d_rays, len(rays), any_hit=any_hit)
        ctx.free(d_rays)
        hits, (nodes, tris, insts) = orc.wide_trace(export, rays, any_hit=any_hit)
        assert (got.nodes, got.tris, got.instances, got.rays) == (nodes, tris, insts, len(rays)), key
        if any_hit:
            assert np.array_equal(hits["inst"] != 0xffffffff, ref.trace_any(rays).astype(bool)), key
        else:
            assert H.hits_equal(hits, ref.trace_closest(rays)), key               # wide BVH == binary BVH, bit for bit
            assert H.hits_equal(hits, ctx.trace_closest_host(top, rays)), key     # == the timed GPU kernel


def test_render_counters_equal_the_oracle(scene):
    """The counters of a count_work launch (what bench.py turns into algorithmic bytes) for the PRIMARY rays of one iteration."""
    name, app, ref = scene
    ctx = app.context(0)
    sysd = app.system_data(0)
    w, h = app.resolution
    app.render(1)
    app.synchronize()
    sysd = app.system_data(0)
    saved = sysd.pathLengths.y
    sysd.pathLengths.y = 1                      # one segment: extend traces exactly the primary rays
    ctx.launch_counts_reset()
    ctx.launch_ex(sysd, w, h, core.RAYGEN_FULL_FRAME, app.info.miss, 5, 1, 0, True)
    ctx.synchronize()
    ext, _ = ctx.launch_counts()
    sysd.pathLengths.y = saved
    primary = ref.generate_primary(H.oracle_sys(app), w, h, 5)
    _, (nodes, tris, insts) = orc.wide_trace(ctx.scene_export(sysd.topObject), primary[primary["tmax"] > 0])
    assert (ext.nodes, ext.tris, ext.instances, ext.rays) == (nodes, tris, insts, int((primary["tmax"] > 0).sum()))
