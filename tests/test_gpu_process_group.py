"""GPU parity, part 5: one process per GPU (Raytracer::joinProcessGroup) -- the sample-range partition with the frames
combined by ncclReduce(mean) inside the host library.

  * world = 1 on one GPU: joining a group of one must not change a single bit of the frame.
  * world = 2 on two GPUs (skipped with fewer): two processes, each rank renders its half of the iteration indices; rank 0's
    collective frame must be bit-identical to 0.5 * (oracle average of iterations 0..S-1 + oracle average of S..2S-1),
    and statistically the single-GPU image of all 2S iterations.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import helpers as H
from tweeker_raytracer_b200 import core, host

pytestmark = pytest.mark.gpu

W, HGT, SQRT = 96, 64, 4      # 16 spp in total


def _system(tmp, device=0):
    return H.write_system(tmp, "rtigo3_cornell_box", resolution="%d %d" % (W, HGT), samplesSqrt=SQRT, devicesMask=1 << device, strategy=0)


def test_group_of_one_is_identity(cuda_device, tmp_path):
    with host.App(_system(tmp_path), H.scene_path("rtigo3_cornell_box")) as app:
        app.render(SQRT * SQRT)
        plain = app.frame()
    with host.App(_system(tmp_path), H.scene_path("rtigo3_cornell_box")) as app:
        app.join_group(0, 1, host.process_group_id())
        assert app.render(5) == 5
        assert app.render(100) == SQRT * SQRT        # the budget of a group of one is the whole sample count
        joined = app.frame()
    assert joined.tobytes() == plain.tobytes()


_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import helpers as H
from tweeker_raytracer_b200 import host
rank, world, tmp = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
idfile = os.path.join(tmp, "nccl_id")
if rank == 0:
    with open(idfile + ".tmp", "wb") as f:
        f.write(host.process_group_id())
    os.rename(idfile + ".tmp", idfile)
else:
    import time
    while not os.path.exists(idfile):
        time.sleep(0.05)
gid = open(idfile, "rb").read()
system = H.write_system(os.path.join(tmp, "rank%d" % rank), "rtigo3_cornell_box", resolution="{w} {h}", samplesSqrt={sqrt}, devicesMask=1 << rank, strategy=0)
with host.App(system, H.scene_path("rtigo3_cornell_box")) as app:
    app.join_group(rank, world, gid)
    budget = {sqrt} * {sqrt} // world
    assert app.render(3) == 3
    assert app.render(1000) == budget          # stops at this rank's share
    frame = app.frame()                        # collective: the mean frame on rank 0, None elsewhere
    assert (frame is None) == (rank != 0)
    np.save(os.path.join(tmp, "frame%d.npy" % rank), frame if rank == 0 else app.local_frame())
"""


def test_two_processes_two_gpus(cuda_device, tmp_path):
    if cuda_device < 2:
        pytest.skip("needs two GPUs")
    tmp = str(tmp_path)
    for r in range(2):
        os.makedirs(os.path.join(tmp, "rank%d" % r))
    script = os.path.join(tmp, "worker.py")
    with open(script, "w") as f:
        f.write(_WORKER.format(root=H.ROOT, w=W, h=HGT, sqrt=SQRT))
    procs = [subprocess.Popen([sys.executable, script, str(r), "2", tmp]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    got = np.load(os.path.join(tmp, "frame0.npy"))
    local1 = np.load(os.path.join(tmp, "frame1.npy"))

    app = host.App(_system(tmp_path), H.scene_path("rtigo3_cornell_box"), host_only=True)
    ref, sysd = H.oracle_scene(app), H.oracle_sys(app)
    half = SQRT * SQRT // 2
    xy = np.array([(x, y) for y in range(HGT) for x in range(W)], dtype=np.uint32)

    def average(first):
        frame = np.zeros((W * HGT, 4), dtype=np.float32)
        for k in range(half):
            one = ref.path_radiance(sysd, app.info.miss, W, xy, first + k)
            frame[:, :3] = one[:, :3] if k == 0 else frame[:, :3] + np.float32(1.0 / (k + 1)) * (one[:, :3] - frame[:, :3])
            frame[:, 3] = 1.0
        return frame.reshape(HGT, W, 4)

    a0, a1 = average(0), average(half)
    assert local1.tobytes() == a1.tobytes()                       # local_frame() of rank 1: its own running average
    want = (a0 * np.float32(0.5) + a1 * np.float32(0.5)).astype(np.float32)
    assert got.tobytes() == want.tobytes()                        # mean of two: exact in binary floating point
    single = ref.render(sysd, app.info.miss, W, HGT, iter_count=2 * half).reshape(HGT, W, 4)
    app.close()
    assert np.allclose(got[..., :3], single[..., :3], rtol=1e-4, atol=1e-5)
