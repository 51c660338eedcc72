#!/usr/bin/env python
"""BASELINE config 5 with the TILE partition (reference strategy 3, local copy; one process driving N GPUs; tiles combined by
one ncclReduce, `composite 1`): 3840x2160, 1024 spp.  `rtigo3_b200 -m 1` times a cold process like the reference's benchmark
mode does (allocations and the first launches included); this script adds the steady state: one untimed 32-spp warm-up, a
restart, then the timed 1024 spp including the composite and the read-back of the frame.
usage: tools/tile_partition_bench.py [num_gpus]"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from tweeker_raytracer_b200 import host


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    tmp = tempfile.mkdtemp()
    system = H.write_system(tmp, "rtigo3_geometry_4k_tiles", devicesMask=(1 << n) - 1)
    app = host.App(system, H.scene_path("rtigo3_geometry"))
    w, h = app.resolution
    spp = app.spp
    t0 = time.perf_counter()
    done = 0
    while done < spp:
        done = app.render(32)
    app.frame_view()
    cold = time.perf_counter() - t0
    app.restart()
    t0 = time.perf_counter()
    done = 0
    while done < spp:
        done = app.render(32)
    frame = app.frame_view()
    warm = time.perf_counter() - t0
    print("tile partition, %d GPUs, %dx%d, %d spp: cold %.3f s = %.2f G samples/s, steady state %.3f s = %.2f G samples/s (frame mean %.4f)"
          % (n, w, h, spp, cold, w * h * spp / cold / 1e9, warm, w * h * spp / warm / 1e9, float(frame[..., :3].mean())))
    app.close()


if __name__ == "__main__":
    main()
