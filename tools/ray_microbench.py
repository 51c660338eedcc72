#!/usr/bin/env python
"""BASELINE config 3: ray microbenchmark on synthetic triangle soups.

Soups of N triangles (centres uniform in the unit cube, edge ~ N^(-1/3)) are built by the device LBVH builder; R rays are
traced as closest-hit and as any-hit (tmax 0.5), coherent (pinhole primaries) and incoherent (uniform origins and
directions).  Reports Mrays/s (CUDA events on the context stream), the build time and the algorithmic bytes per ray
(48 + 80 nodes + 48 triangles + 64 instances, counted by rtc_trace_count on a 1/16 subset) with the resulting GB/s.
torch is used only to synthesise the soup and the rays on the device.

  python tools/ray_microbench.py --tris 1e6,1e7 --rays 1e8 [--out gpurun_out/microbench.jsonl]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tweeker_raytracer_b200 import core  # noqa: E402


def make_soup(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    edge = float(n) ** (-1.0 / 3.0)
    c = torch.rand((n, 1, 3), device="cuda", generator=g)
    v = (c + (torch.rand((n, 3, 3), device="cuda", generator=g) - 0.5) * (2.0 * edge)).reshape(-1, 3).contiguous()
    idx = torch.arange(3 * n, device="cuda", dtype=torch.int32)
    return v, idx


def coherent_rays(n, tmax):
    side = int(n ** 0.5)
    n = side * side
    ys, xs = torch.meshgrid(torch.arange(side, device="cuda"), torch.arange(side, device="cuda"), indexing="ij")
    ndc = torch.stack([(xs.reshape(-1) + 0.5) / side * 2 - 1, (ys.reshape(-1) + 0.5) / side * 2 - 1], dim=1)
    d = torch.stack([ndc[:, 0] * 0.6, ndc[:, 1] * 0.6, -torch.ones(n, device="cuda")], dim=1)
    d = d / d.norm(dim=1, keepdim=True)
    rays = torch.empty((n, 8), device="cuda", dtype=torch.float32)
    rays[:, 0], rays[:, 1], rays[:, 2], rays[:, 3] = 0.5, 0.5, 2.2, 1e-5
    rays[:, 4:7] = d
    rays[:, 7] = tmax
    return rays


def incoherent_rays(n, tmax, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    rays = torch.empty((n, 8), device="cuda", dtype=torch.float32)
    rays[:, 0:3] = torch.rand((n, 3), device="cuda", generator=g)
    rays[:, 3] = 1e-5
    d = torch.randn((n, 3), device="cuda", generator=g)
    rays[:, 4:7] = d / d.norm(dim=1, keepdim=True)
    rays[:, 7] = tmax
    return rays


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tris", default="1e6,1e7")
    ap.add_argument("--rays", default="1e8")
    ap.add_argument("--out", default=None)
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    nrays = int(float(args.rays))
    peak = 6548.2
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    ctx = core.Context(0)
    lines = []
    for nt in [int(float(t)) for t in args.tris.split(",")]:
        verts, idx = make_soup(nt, 0x1234567)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gas = ctx.gas_build(verts.data_ptr(), 12, 3 * nt, idx.data_ptr(), nt, core.BUILD_GPU_LBVH)
        ctx.synchronize()
        first_build_s = time.perf_counter() - t0          # includes first-use cudaMalloc / module load effects
        ctx.gas_destroy(gas)
        t0 = time.perf_counter()
        gas = ctx.gas_build(verts.data_ptr(), 12, 3 * nt, idx.data_ptr(), nt, core.BUILD_GPU_LBVH)
        ctx.synchronize()
        build_s = time.perf_counter() - t0                # steady state: scratch allocation + kernels + the per-level host syncs
        inst = __import__("numpy").zeros(1, dtype=core.INSTANCE_DTYPE)
        inst[0]["transform"] = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]
        inst[0]["gas"] = gas
        top = ctx.ias_build(inst)
        info = ctx.scene_info(top)
        for kind in ("coherent", "incoherent"):
            for mode, tmax in (("closest", 1e27), ("any", 1e27 if kind == "coherent" else 0.5)):
                rays = coherent_rays(nrays, tmax) if kind == "coherent" else incoherent_rays(nrays, tmax, 0x89ABCDEF)
                n = rays.shape[0]
                out = torch.empty(n * (5 if mode == "closest" else 1), device="cuda", dtype=torch.int32)
                torch.cuda.synchronize()
                best = None
                for _ in range(args.repeat + 1):          # first pass is the warm-up
                    ctx.timer_start()
                    if mode == "closest":
                        ctx.trace_closest(top, rays.data_ptr(), n, out.data_ptr())
                    else:
                        ctx.trace_any(top, rays.data_ptr(), n, out.data_ptr())
                    ms = ctx.timer_stop()
                    best = ms if best is None or _ == 1 else min(best, ms)
                subset = rays[::16].contiguous()          # every 16th ray: same distribution as the full set
                torch.cuda.synchronize()                  # torch's stream and the context's stream are not ordered
                counts = ctx.trace_count(top, subset.data_ptr(), subset.shape[0], any_hit=(mode == "any"))
                per_ray = (48 * counts.rays + 80 * counts.nodes + 48 * counts.tris + 64 * counts.instances) / max(counts.rays, 1)
                if mode == "closest":
                    hit_rate = float((out.view(n, 5)[:, 3] != -1).float().mean().item())
                else:
                    hit_rate = float(out.float().mean().item())
                line = {"triangles": nt, "rays": n, "kind": kind, "mode": mode, "mrays_per_s": n / (best * 1e-3) / 1e6, "ms": best,
                        "hit_rate": hit_rate, "nodes_per_ray": counts.nodes / max(counts.rays, 1), "tris_per_ray": counts.tris / max(counts.rays, 1),
                        "algorithmic_bytes_per_ray": per_ray, "achieved_gbs": per_ray * n / (best * 1e-3) / 1e9,
                        "frac_of_measured_hbm": per_ray * n / (best * 1e-3) / 1e9 / peak,
                        "build_s": build_s, "first_build_s": first_build_s, "build_mtris_per_s": nt / build_s / 1e6, "bvh_nodes": int(info.numNodes),
                        "bvh_mb": (info.numNodes * 80 + info.numTris * 48) / 1e6}
                lines.append(line)
                print(json.dumps(line), flush=True)
                del rays, out, subset
        ctx.scene_destroy(top)
        ctx.gas_destroy(gas)
        del verts, idx
        torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as f:
            for l in lines:
                f.write(json.dumps(l) + "\n")
    ctx.close()


if __name__ == "__main__":
    main()
