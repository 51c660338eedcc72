#!/usr/bin/env python
"""Prints how the ray pool of the traversal kernels spent its passes on one step of a scene (count_work launch):
per phase the warp-level passes, the slots they processed and the mean lanes per pass, plus the roofline probes.
usage: tools/pool_stats.py [scene] [resolution] [spp]"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from tweeker_raytracer_b200 import core, host


def main():
    scene = sys.argv[1] if len(sys.argv) > 1 else "rtigo3_geometry"
    res = sys.argv[2] if len(sys.argv) > 2 else "1920 1080"
    spp = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    tmp = tempfile.mkdtemp()
    app = host.App(H.write_system(tmp, scene, resolution=res, samplesSqrt=16, devicesMask=1, strategy=0), H.scene_path(scene))
    w, h = app.resolution
    ctx = app.context(0)
    app.render(1)
    app.synchronize()
    sysd = app.system_data(0)
    ctx.launch_counts_reset()
    ctx.launch_ex(sysd, w, h, core.RAYGEN_FULL_FRAME, app.info.miss, 0, spp, 0, True)
    ctx.synchronize()
    for name, c, ps in zip(("extend", "connect"), ctx.launch_counts(), ctx.launch_pass_stats()):
        r = max(c.rays, 1)
        print("%s: rays %d  nodes/ray %.3f tris/ray %.3f insts/ray %.3f" % (name, c.rays, c.nodes / r, c.tris / r, c.instances / r))
        tot = sum(v[0] for v in ps.values())
        for k, (passes, lanes, mean) in ps.items():
            print("   %-9s passes %10d (%5.1f%%)  slots %11d  lanes/pass %5.2f" % (k, passes, 100.0 * passes / max(tot, 1), lanes, mean))
    print("probe: L2 gather (32 MB set) %.0f GB/s, HBM gather (8 GB set) %.0f GB/s, fp32 %.1f TFLOP/s, issue %.0f G warp-inst/s" % (
        ctx.probe_gather(32 << 20), ctx.probe_gather(8 << 30), ctx.probe_pipes(0), ctx.probe_pipes(1)))
    app.close()


if __name__ == "__main__":
    main()
