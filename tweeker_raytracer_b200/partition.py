"""Work partition of one render across GPUs (one process per GPU).  Pure index arithmetic, no device code.

Two partitions of the same job:

  * tile partition  -- the reference's strategies 1-3: a checkerboard of tileSize tiles, row yBlock rotated by the device
    index (apps/rtigo3/shaders/raygeneration.cu:152-164); each device launches `launch_width` columns
    (apps/rtigo3/src/DeviceMultiGPULocalCopy.cpp:91-93).  Seeds depend on the device layout.
  * sample-range partition -- the samplesPerPixel iterations of the render are cut into n contiguous ranges; device r
    renders iteration indices [r*spp/n, (r+1)*spp/n) over the whole frame and keeps its OWN running average (accumulation
    index counts from 0); the frame is the mean of the n averages, obtained with one NCCL reduce over NVLink.  Seeds are
    the single-GPU ones: the n ranks together draw exactly the samples of the 1-GPU render.  Mirrors
    Raytracer::samplesPerRank / joinProcessGroup (host/Raytracer.h); host.sample_range is the C++ side of it.
"""


def tiled_launch_width(resolution_x, device_count, tile_size_x):
    width = (resolution_x + device_count - 1) // device_count
    mask = tile_size_x - 1
    return (width + mask) & ~mask


def tile_shift(tile_size):
    shift = 0
    while shift < 32 and (tile_size & (1 << shift)) == 0:
        shift += 1
    return shift


def distribute(x, y, device_index, device_count, tile_size_x, tile_shift_x, tile_shift_y):
    """Launch index (x, y) of one device -> pixel column (raygeneration.cu:152-164)."""
    x_block = x >> tile_shift_x
    y_block = y >> tile_shift_y
    x_tile = x_block * device_count + ((device_index + y_block) % device_count)
    return x_tile * tile_size_x + (x & (tile_size_x - 1))


def samples_per_rank(samples_per_pixel, world):
    """Iterations each rank renders out of a budget of samples_per_pixel (at least 1)."""
    return max(samples_per_pixel // max(world, 1), 1)


def sample_range(step, rank, world, spp_per_step, samples_per_pixel):
    """(first seed iteration, count, first accumulation index) of `rank` in `step` (a step = spp_per_step iterations)."""
    local = samples_per_rank(samples_per_pixel, world)
    offset = rank * local if world > 1 else 0
    first = step * spp_per_step
    count = max(0, min(spp_per_step, local - first))
    return offset + first, count, first


def combine_scale(world):
    """Factor applied after the sum-reduce of the per-rank running averages."""
    return 1.0 / world


def bench_step_range(step, rank, world, spp_per_step, scaling="weak"):
    """(first seed iteration, count, first accumulation index) of `rank` in bench.py's timed step `step`.

    weak  : every rank renders spp_per_step iterations per step (total work grows with the number of GPUs);
    strong: the step's spp_per_step iterations are split, spp_per_step / world per rank (total work is fixed), so the
            ranks of step s together draw exactly the seed iterations [s * spp_per_step, (s + 1) * spp_per_step) of the
            one-GPU step -- the combined frame holds the same samples whatever the number of GPUs.
    Either way the ranks' iteration indices are disjoint and each rank accumulates its own running average from index 0."""
    if scaling == "strong":
        if spp_per_step % world != 0:
            raise ValueError("strong scaling needs spp_per_step divisible by the number of ranks")
        count = spp_per_step // world
    else:
        count = spp_per_step
    return (step * world + rank) * count, count, step * count
