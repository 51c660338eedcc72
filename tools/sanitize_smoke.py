#!/usr/bin/env python
"""Small render + ray queries for compute-sanitizer (memcheck / racecheck / initcheck): the Cornell box at 64x48, 4 spp through
Application -> Raytracer -> Device -> librtcore, the geometry scene with the HDR environment at 48x32, the textured scene with
the ordered any-hit rounds at 48x32, and 20 000 incoherent ray queries -- every kernel of the hot path, including the queue
compaction (block_append*).  usage: compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from tweeker_raytracer_b200 import host


def main():
    tmp = tempfile.mkdtemp()
    for name, kw in (("rtigo3_cornell_box", dict(resolution="64 48", samplesSqrt=2)),
                     ("rtigo3_geometry", dict(resolution="48 32", samplesSqrt=2, miss=2, envMap="procedural 64 32")),
                     ("rtigo3_textures", dict(resolution="48 32", samplesSqrt=2))):
        app = host.App(H.write_system(tmp, name, **kw), H.scene_path(name))
        app.render_calls(3)
        app.render(1)
        frame = app.frame()
        ctx = app.context(0)
        top = app.system_data(0).topObject
        rays = H.random_rays(20000, seed=11, lo=(-2, 0, -2), hi=(2, 3, 2))
        hits = ctx.trace_closest_host(top, rays)
        occ = ctx.trace_any_host(top, rays)
        st = app.stats()
        print("%s: frame mean %.4f, %d of %d query rays hit, %d occluded, %d launches, stack overflows %d"
              % (name, float(frame[..., :3].mean()), int((hits["inst"] != 0xffffffff).sum()), len(rays), int(occ.sum()), st.kernelLaunches, st.stackOverflows))
        app.close()


if __name__ == "__main__":
    main()
