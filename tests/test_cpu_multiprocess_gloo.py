"""world_size-2 test of the sample-range partition on CPU (gloo): each rank renders ITS iteration indices with the
oracle, keeps its own running average, the frames are combined with one reduce(sum) and the 1/n scale -- the same
arithmetic bench.py runs over NCCL.  The result must equal the mean of the per-rank averages computed in one process and
be statistically the single-process image."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def rank_average(tmp, rank, world, steps, spp):
    from tweeker_raytracer_b200 import host, partition
    app = host.App(H.write_system(tmp, "rtigo3_cornell_box", resolution="16 16"), H.scene_path("rtigo3_cornell_box"), host_only=True)
    scene, sysd = H.oracle_scene(app), H.oracle_sys(app)
    frame = np.zeros((256, 4), dtype=np.float32)
    xy = np.array([(x, y) for y in range(16) for x in range(16)], dtype=np.uint32)
    for step in range(steps):
        first, count, accum = partition.sample_range(step, rank, world, spp, steps * spp * world)
        # the oracle blends sample `it` with weight 1/(it+1); emulate accumulation index != seed index one iteration at a time
        for k in range(count):
            one = scene.path_radiance(sysd, 0, 16, xy, first + k)      # raw radiance of seed iteration first + k
            n = accum + k
            frame[:, :3] = one[:, :3] if n == 0 else frame[:, :3] + np.float32(1.0 / (n + 1)) * (one[:, :3] - frame[:, :3])
            frame[:, 3] = 1.0
    app.close()
    return frame


def _worker(rank, world, port, tmp, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tweeker_raytracer_b200 import partition
    frame = torch.from_numpy(rank_average(tmp, rank, world, steps=2, spp=2))
    dist.reduce(frame, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        frame *= partition.combine_scale(world)
        np.save(out, frame.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sample_range_partition_world2(built, tmp_path):
    out = os.path.join(str(tmp_path), "combined.npy")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), out), nprocs=2, join=True)
    got = np.load(out)
    parts = [rank_average(str(tmp_path), r, 2, steps=2, spp=2) for r in range(2)]
    want = (parts[0] + parts[1]) * np.float32(0.5)
    assert np.array_equal(got, want)
    assert np.allclose(got[:, 3], 1.0)
    # the two ranks drew disjoint iteration indices 0..7: together they are the 8-spp single-process image up to rounding
    from tweeker_raytracer_b200 import host
    app = host.App(H.write_system(str(tmp_path), "rtigo3_cornell_box", resolution="16 16"), H.scene_path("rtigo3_cornell_box"), host_only=True)
    single = H.oracle_scene(app).render(H.oracle_sys(app), 0, 16, 16, iter_count=8)
    app.close()
    assert np.allclose(got[:, :3], single[:, :3], rtol=1e-4, atol=1e-5)


def test_sample_range_matches_host_library(built):
    """partition.sample_range (Python) == Raytracer::samplesPerRank / seed offsets (C++ host), and the ranges tile the budget."""
    from tweeker_raytracer_b200 import host, partition
    for spp, world in [(256, 1), (256, 2), (256, 8), (1024, 4), (16, 8), (4, 8), (100, 3)]:
        covered = []
        for rank in range(world):
            first, count = host.sample_range(spp, rank, world)
            assert count == partition.samples_per_rank(spp, world)
            assert (first, count, 0) == partition.sample_range(0, rank, world, count, spp)
            # stepping through the range in chunks of 3 visits it exactly once
            seen, step = [], 0
            while True:
                f, c, a = partition.sample_range(step, rank, world, 3, spp)
                if c == 0:
                    break
                assert f == first + a
                seen.extend(range(f, f + c))
                step += 1
            assert seen == list(range(first, first + count))
            covered.extend(seen)
        assert len(set(covered)) == len(covered)
        if spp % world == 0:
            assert sorted(covered) == list(range(spp))


def test_bench_step_ranges_weak_and_strong():
    """bench.py's partition of a timed step: disjoint iteration indices per rank; under strong scaling the ranks of a step
    together draw exactly the one-GPU step's seed iterations, so the combined frame holds the same samples for every N."""
    from tweeker_raytracer_b200 import partition
    for world in (1, 2, 4, 8):
        for scaling in ("weak", "strong"):
            seen = []
            for step in range(3):
                per_step = []
                for rank in range(world):
                    first, count, accum = partition.bench_step_range(step, rank, world, 32, scaling)
                    assert count == (32 // world if scaling == "strong" else 32) and accum == step * count
                    per_step += list(range(first, first + count))
                if scaling == "strong":
                    assert sorted(per_step) == list(range(step * 32, (step + 1) * 32))
                seen += per_step
            assert len(seen) == len(set(seen))
    with pytest.raises(ValueError):
        partition.bench_step_range(0, 0, 3, 32, "strong")
