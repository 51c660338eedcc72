#!/usr/bin/env python
"""bench.py -- the headline measurement: 1080p samples/s (and Mrays/s) of the rtigo3 geometry scene.

  python bench.py --gpus N --steps K --warmup W            the B200 core (this repository)
  python bench.py --impl reference --gpus N ...            the reference's own device programs compiled for the host
                                                           (oracle/_ref, one process per core; oracle port if absent)

A "step" renders `--spp-per-step` iterations (samples per pixel) of the 1920x1080 frame: generate -> [extend -> shade ->
connect] x depth -> accumulate, i.e. one pass of the hot path over one batch of 1920*1080*spp path samples.

  value     whole-job Msamples/s (pixel-samples per second / 1e6) with the scene resident in HBM, device-timed with CUDA
            events on the launching stream, max over ranks.  N > 1: sample-range partition (each GPU renders its own
            iteration indices over the full frame, scaling "weak") followed by ONE ncclReduce(mean) of the accumulation
            buffers over NVLink on the render stream (the host library's own communicator), inside the timed region.
  e2e       the same metric through the reference-facing classes (Application::render -> Raytracer -> Device ->
            librtcore) with the camera uploaded from host memory and the float4 frame read back to host memory every step.
  roofline  the extend (closest-hit traversal) kernel: algorithmic bytes (rays 48 B + nodes 80 B + triangles 48 B +
            instance records 64 B, counted by a second, untimed pass with the same seeds) / its device time.
  cpu_baseline  oracle/_ref (the reference's shader sources host-compiled, kind "reference"; traversal served by the
            oracle's intersector) on a bounded sample of the same workload, one process per host core; kind "port"
            (oracle/rt_oracle.c, threads) where libref.so is not available.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "rtigo3_geometry_1080p_samples_per_s"
UNIT = "Msamples/s"
WORKLOAD = "rtigo3 geometry scene (planes/boxes/spheres/tori, 5 BSDFs, constant env + 4x4 parallelogram light), pathLengths 2 6"
S_RAY, S_NODE, S_TRI, S_INST = 48, 80, 48, 64
WORKLOADS = {
    "rtigo3_geometry": WORKLOAD,
    "rtigo3_cornell_box": "rtigo3 Cornell box (area light, mirror + glass spheres)",
    "rtigo3_instances": "instanced stress scene: instances of a 50 000-triangle torus (two-level BVH), constant environment",
    "rtigo3_textures": "rtigo3 geometry scene with albedo and cutout textures (ordered any-hit processing, host-synchronised rounds)",
}


def workload(args):
    return "%s, %s" % (WORKLOADS.get(args.scene, args.scene), args.resolution.replace(" ", "x"))


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spp-per-step", type=int, default=32)
    ap.add_argument("--resolution", default="1920 1080")
    ap.add_argument("--scene", default="rtigo3_geometry")
    ap.add_argument("--instances", type=int, default=10000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clocks, power and clock-event (throttle) reasons of ONE GPU while the timed region runs.

    In-process NVML (pynvml) on a handle looked up once by UUID: ~1 kHz capable, sampled every 5 ms, and it does not
    spawn a process per sample (nvidia-smi attaches to every GPU of the box on each call, which N ranks polling at once
    would feel).  Falls back to polling nvidia-smi where pynvml is missing."""

    def __init__(self, index, uuid=None):
        super().__init__(daemon=True)
        self.index = index
        self.uuid = uuid
        self.stop_flag = threading.Event()
        self.rows = []          # (sm_mhz, sm_max_mhz, power_w, reason_bits)
        self.source = "nvml"
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = (pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode()) if uuid
                           else pynvml.nvmlDeviceGetHandleByIndex(index))
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.source = "nvidia-smi"

    def _sample_nvml(self):
        nv = self.nvml
        sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
        try:
            bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        try:
            power = nv.nvmlDeviceGetPowerUsage(self.handle) / 1e3
        except Exception:
            power = None
        reasons = set()
        for name, bit in (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap),
                          ("hw_power_brake_slowdown", nv.nvmlClocksEventReasonHwPowerBrakeSlowdown)):
            if bits & bit:
                reasons.add(name)
        self.rows.append((sm, self.sm_max, power, reasons))

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if not out:
            return
        r = [c.strip() for c in out.splitlines()[0].split(",")]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = {name for k, name in enumerate(names) if len(r) > 3 + k and r[3 + k].lower().startswith("active")}
        self.rows.append((float(r[0]), float(r[1]), float(r[2]) if r[2].replace(".", "").isdigit() else None, reasons))

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.005 if self.nvml is not None else 0.05)

    def summary(self):
        sm = sorted(r[0] for r in self.rows)
        power = [r[2] for r in self.rows if r[2] is not None]
        reasons = set()
        for r in self.rows:
            reasons |= r[3]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": max(r[1] for r in self.rows) if self.rows else None, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(self.rows), "source": self.source}


def measured_traffic():
    """DRAM bytes per extend launch from the committed ncu capture (profiles/extend_traffic_r1.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "extend_traffic_r1.json")) as f:
            d = json.load(f)
        return float(d["traffic_bytes_per_launch"]), d["source"]
    except Exception:
        return None, None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def system_file(tmp, args, device_ordinal):
    import helpers as H
    return H.write_system(tmp, args.scene, resolution=args.resolution, samplesSqrt=256, devicesMask=1 << device_ordinal, strategy=0)


def scene_file(tmp, args):
    """scenes/scene_<name>.txt; the instanced stress scene (config 4) is generated (tools/make_instances_scene.py)."""
    import helpers as H
    if args.scene == "rtigo3_instances":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import make_instances_scene
        path = os.path.join(tmp, "scene_rtigo3_instances.txt")
        if not os.path.exists(path):
            make_instances_scene.write_scene(path, count=args.instances)
        return path
    return H.scene_path(args.scene)


def cpu_sample(args, threads, iterations=2, row_step=16, host_only_app=None):
    """Oracle (kind "port") on rows y % row_step == 0 for `iterations` samples per pixel; returns (Msamples/s, seconds, description, Mrays/s)."""
    import helpers as H
    from oracle import orc
    from tweeker_raytracer_b200 import host
    tmp = tempfile.mkdtemp()
    app = host_only_app or host.App(system_file(tmp, args, 0), scene_file(tmp, args), host_only=True)
    ref = H.oracle_scene(app)
    w, h = app.resolution
    sysd = H.oracle_sys(app)
    st = orc.Stats()
    t0 = time.perf_counter()
    ref.render(sysd, app.info.miss, w, h, iter_count=iterations, row_step=row_step, row_offset=0, threads=threads, stats=st)
    dt = time.perf_counter() - t0
    desc = "rows y%%%d==0 of %dx%d, %d spp = %d path samples (%d radiance + %d shadow rays) in %.2f s" % (
        row_step, w, h, iterations, st.pathSamples, st.radianceRays, st.shadowRays, dt)
    return st.pathSamples / dt / 1e6, dt, desc, (st.radianceRays + st.shadowRays) / dt / 1e6


# ---- the reference's own device programs, compiled for the host (oracle/_ref/libref.so), one process per core -------------
_REF = {}


def _ref_init(system_path, scene_path):
    import helpers as H
    from oracle import orc
    from tweeker_raytracer_b200 import host
    app = host.App(system_path, scene_path, host_only=True)
    scene = H.oracle_scene(app, "libm")
    _REF.update(app=app, scene=scene, ref=orc.Reference(scene, app.info.miss), sysd=H.oracle_sys(app))


def _ref_rows(task):
    k, procs, iterations = task
    app = _REF["app"]
    w, h = app.resolution
    t0 = time.perf_counter()
    _REF["ref"].render(_REF["sysd"], w, h, iter_count=iterations, row_step=procs, row_offset=k)
    return time.perf_counter() - t0


class ReferencePool:
    """The reference keeps its launch parameters in a global, so it is parallelised over processes: worker k renders the
    launch rows y % P == k.  Set-up (scene load, oracle BVH for optixTrace) happens once per worker, outside the timing."""

    def __init__(self, args, procs):
        import multiprocessing as mp
        tmp = tempfile.mkdtemp()
        self.system_path, self.scene_path = system_file(tmp, args, 0), scene_file(tmp, args)
        self.procs = procs
        self.pool = mp.get_context("fork").Pool(procs, initializer=_ref_init, initargs=(self.system_path, self.scene_path))
        self.pool.map(_ref_rows, [(k, 64 * procs, 1) for k in range(procs)])       # touch every worker (set-up done)
        self.resolution = tuple(int(v) for v in args.resolution.split())

    def sample(self, iterations):
        t0 = time.perf_counter()
        self.pool.map(_ref_rows, [(k, self.procs, iterations) for k in range(self.procs)], chunksize=1)
        dt = time.perf_counter() - t0
        w, h = self.resolution
        n = w * h * iterations
        return n / dt / 1e6, dt, "full %dx%d frame, %d spp = %d path samples in %.2f s, %d processes" % (w, h, iterations, n, dt, self.procs)

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_reference(args, iterations, pool=None):
    """(value, seconds, description, cores, kind): oracle/_ref when it is built ("reference"), else the oracle port."""
    from oracle import orc
    cores = orc.online_cores()
    if orc.reference_available():
        own = pool is None
        pool = pool or ReferencePool(args, cores)
        v, dt, desc = pool.sample(iterations)
        if own:
            pool.close()
        return v, dt, desc, cores, "reference"
    v, dt, desc, _ = cpu_sample(args, cores, iterations=iterations, row_step=1)
    return v, dt, desc, cores, "port"


def run_reference(args, rank):
    """The reference arm: the reference's own device programs compiled for the host (oracle/_ref, kind "reference", one
    process per host core; falls back to the oracle port where libref.so does not exist), rank 0 only."""
    if rank != 0:
        return
    from oracle import orc
    cores = orc.online_cores()
    pool = ReferencePool(args, cores) if orc.reference_available() else None
    for _ in range(min(args.warmup, 1)):
        cpu_reference(args, 1, pool)
    vals, secs, desc, kind = [], 0.0, "", "port"
    for _ in range(args.steps):
        v, dt, desc, cores, kind = cpu_reference(args, 16, pool)
        vals.append(v)
        secs += dt
    if pool:
        pool.close()
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": workload(args), "sample_per_step": desc},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": "each step: " + desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import helpers as H
    from tweeker_raytracer_b200 import core, host

    if core.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: " + core.lib().rtc_last_error().decode())
    n = args.gpus
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)
    if world != n and world > 1:
        n = world
    S, K, W = args.spp_per_step, args.steps, args.warmup

    tmp = tempfile.mkdtemp()
    app = host.App(system_file(tmp, args, local_rank), scene_file(tmp, args))
    w, h = app.resolution
    pixels = w * h
    ctx = app.context(0)
    app.render(1)                      # allocates the frame, warms the allocator
    app.synchronize()
    sysd = app.system_data(0)
    info = ctx.scene_info(sysd.topObject)

    # N > 1: the Application joins the process group of the host library (Raytracer::joinProcessGroup: sample-range partition,
    # its own NCCL communicator on the render stream); torch.distributed carries the 128-byte id and the barriers
    if dist is not None:
        ids = [host.process_group_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        app.join_group(rank, world, ids[0])
    # accumulation buffer of the device-timed arm, and the buffer the mean frame lands in on rank 0
    frame = torch.zeros(pixels * 4, dtype=torch.float32, device="cuda")
    combined = torch.zeros(pixels * 4, dtype=torch.float32, device="cuda") if rank == 0 else None
    sysd.outputBuffer = frame.data_ptr()

    def combine():
        # the one exchange step of the path: ncclReduce(mean) of the per-rank running averages over NVLink, on the render stream
        app.group_reduce_mean(frame.data_ptr(), combined.data_ptr() if rank == 0 else 0, pixels * 4)

    def step(s, count_work=False):
        # sample-range partition: rank r of n renders iteration indices (s*n + r)*S .. +S as samples s*S.. of its own average
        ctx.launch_ex(sysd, w, h, core.RAYGEN_FULL_FRAME, app.info.miss, (s * n + rank) * S, S, s * S, count_work)

    def barrier():
        torch.cuda.synchronize()
        ctx.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(W):
        step(s)
    if dist is not None:       # warm the NCCL communicator up (connection set-up is not part of a render)
        combine()
    # NVML is initialised BEFORE the barrier: its start-up time differs from process to process, and anything between the
    # barrier and timer_start shows up as skew in the max-over-ranks time
    sampler = ClockSampler(local_rank, getattr(torch.cuda.get_device_properties(local_rank), "uuid", None))
    ctx.stats_reset()
    ctx.profile_enable(True)
    barrier()
    stamps = [time.time()]
    sampler.start()
    ctx.timer_start()
    stamps.append(time.time())
    for s in range(W, W + K):
        step(s)
    stamps.append(time.time())
    steps_ms = ctx.timer_stop()                 # synchronises the render stream
    stamps.append(time.time())
    reduce_ms = 0.0
    if dist is not None:
        ctx.timer_start()
        combine()
        reduce_ms = ctx.timer_stop()            # on each rank: waiting for the slowest rank + the transfer
    stamps.append(time.time())
    barrier()
    sampler.stop_flag.set()
    sampler.join()
    prof = ctx.profile()
    ctx.profile_enable(False)
    stats = ctx.stats()
    total_ms = steps_ms + reduce_ms
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    rays = torch.tensor([float(stats.radianceRays + stats.shadowRays)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
    total_ms = float(t.item())
    per_rank = [[steps_ms, reduce_ms]]
    if dist is not None:        # diagnostics: each rank's own render time and its wait+reduce time
        g = [torch.zeros(2 + len(stamps), dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(g, torch.tensor([steps_ms, reduce_ms] + stamps, dtype=torch.float64, device="cuda"))
        t00 = min(float(x[2]) for x in g)
        # [render ms, reduce ms, then host wall-clock ms since the first rank left the barrier: barrier exit, timer started,
        #  steps enqueued, render finished, reduce finished]
        per_rank = [[round(float(x[0]), 3), round(float(x[1]), 3)] + [round((float(v) - t00) * 1e3, 3) for v in x[2:]] for x in g]
    value = n * K * S * pixels / (total_ms * 1e-3) / 1e6
    mrays = float(rays.item()) / (total_ms * 1e-3) / 1e6
    launches = int(stats.kernelLaunches)

    # ---- algorithmic bytes of the extend kernel: the same steps again, untimed, with work counters
    ctx.launch_counts_reset()
    for s in range(W, W + K):
        step(s, count_work=True)
    ext, con = ctx.launch_counts()
    ext_bytes = ext.rays * S_RAY + ext.nodes * S_NODE + ext.tris * S_TRI + ext.instances * S_INST
    con_bytes = con.rays * S_RAY + con.nodes * S_NODE + con.tris * S_TRI + con.instances * S_INST
    ext_ms, ext_launches = prof["extend"]
    con_ms, con_launches = prof["connect"]
    peak, peak_src = measured_peak()
    achieved = ext_bytes / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
    traffic, traffic_src = measured_traffic() if args.scene == "rtigo3_geometry" else (None, None)
    roofline = {"bound": "hbm", "kernel": "k_trace<ANY=0, ExtendPaths> (closest-hit traversal of the radiance-ray queue)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": ext_bytes / max(ext_launches, 1),
                "avg_launch_ms": ext_ms / max(ext_launches, 1), "launches": int(ext_launches),
                "per_ray": {"nodes": ext.nodes / max(ext.rays, 1), "tris": ext.tris / max(ext.rays, 1), "instances": ext.instances / max(ext.rays, 1),
                            "bytes": ext_bytes / max(ext.rays, 1)},
                "extend_mrays_per_s": ext.rays / (ext_ms * 1e-3) / 1e6 if ext_ms > 0 else 0.0,
                "connect": {"achieved": con_bytes / (con_ms * 1e-3) / 1e9 if con_ms > 0 else 0.0,
                            "mrays_per_s": con.rays / (con_ms * 1e-3) / 1e6 if con_ms > 0 else 0.0,
                            "bytes_per_ray": con_bytes / max(con.rays, 1)},
                "kernel_share_of_step": {k: v[0] / max(sum(x[0] for x in prof.values()), 1e-9) for k, v in prof.items()},
                "note": "BVH (%.1f MB) + triangles fit the 126 MB L2, so DRAM traffic is far below the algorithmic bytes; HBM peak is the contract's denominator"
                        % ((info.numNodes * 80 + info.numTris * 48) / 1e6)}

    # ---- end to end through Application::render with host buffers.  N > 1: the Application joins the process group
    # (sample-range partition inside the host library, its own NCCL communicator), every rank renders its range and
    # fetching the frame is a collective: ncclReduce(mean) over NVLink to rank 0, which reads the combined frame back.
    app.restart()
    cam = app.camera()
    pinned = ctx.host_alloc(48)
    import ctypes
    ctypes.memmove(pinned, cam.ctypes.data, 48)
    sys_host = app.system_data(0)
    for _ in range(min(W, 3)):
        app.render(S)
        app.frame_view()
    barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for _ in range(K):
        ctx.upload_async(sys_host.cameraDefinitions, pinned, 48)      # this step's camera, from pinned host memory
        app.render(S)
        fr = app.frame_view()                                         # device -> host read of the step's result (rank 0 in a group)
        if fr is not None:
            checksum += float(fr[0, 0, 0])
    ctx.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n * K * S * pixels / float(te.item()) / 1e6
    ctx.host_free(pinned)

    line = None
    if rank == 0:
        line = {"metric": METRIC if args.scene == "rtigo3_geometry" else METRIC.replace("rtigo3_geometry", args.scene), "value": value, "unit": UNIT, "n_gpus": n, "steps": K, "warmup": W, "ms_per_step": total_ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload(args), "spp_per_step": S, "path_samples_per_step_per_gpu": S * pixels,
                           "parallelism": "sample-range x%d + NCCL reduce" % n if n > 1 else "single GPU",
                           "l2": "wavefront state per step (%.0f MB) exceeds L2 (126 MB); no explicit flush" % (S * pixels * 292 / 1e6),
                           "triangles": int(info.numTris), "bvh_nodes": int(info.numNodes), "instances": int(info.numInstances)},
                "mrays_per_s": mrays, "reduce_ms": reduce_ms, "per_rank_render_ms_and_reduce_ms": per_rank,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 48 + 192, "d2h_bytes_per_step": pixels * 16,
                        "path": "Application::render + getOutputBufferHost per step" + (" (collective: ncclReduce mean to rank 0, then read back)" if n > 1 else "")},
                "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roofline}
        if n == 1 and not args.no_cpu_baseline:
            v, dt, desc, cores, kind = cpu_reference(args, 96)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}
        print(json.dumps(line), flush=True)
    app.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
