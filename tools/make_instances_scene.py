#!/usr/bin/env python
"""BASELINE config 4: writes the instanced stress scene in rtigo3's scene description format:
`count` instances of ONE tessellated torus (`model torus 250 100 0.75 0.25` = 50 000 triangles, shared through the
loader's geometry cache, apps/rtigo3/src/Application.cpp:1815-1833) on a lattice, with LCG-jittered rotations, over a floor.

  python tools/make_instances_scene.py [--count 10000] [--tess 250 100] out_scene.txt
"""
import argparse


def lcg(state):
    state = (state * 1664525 + 1013904223) & 0xffffffff
    return state, (state & 0x00ffffff) / float(0x01000000)


def write_scene(path, count=10000, tess=(250, 100), spacing=2.4):
    side = int(round(count ** (1.0 / 3.0)))
    nx = nz = max(1, int(round((count / max(side // 2, 1)) ** 0.5)))
    ny = max(1, (count + nx * nz - 1) // (nx * nz))
    lines = ["# instanced stress scene: %d instances of torus %d x %d (%d triangles each, one shared GAS)" % (count, tess[0], tess[1], 2 * tess[0] * tess[1]),
             "albedo 0.5 0.5 0.5", "material floor brdf_diffuse", "material default brdf_diffuse"]
    palette = [("red", "0.8 0.15 0.1", "brdf_diffuse"), ("green", "0.15 0.7 0.2", "brdf_diffuse"), ("blue", "0.15 0.25 0.8", "brdf_diffuse"),
               ("gold", "0.9 0.7 0.3", "brdf_ggx_smith"), ("mirror", "0.95 0.95 0.95", "brdf_specular"), ("glass", "1 1 1", "bsdf_specular")]
    for name, albedo, bsdf in palette:
        lines += ["albedo " + albedo, "roughness 0.15 0.15", "material %s %s" % (name, bsdf)]
    lines += ["identity", "push scale %g 1 %g model plane 16 16 1 floor pop" % (nx * spacing, nz * spacing)]
    state = 0x1234567
    placed = 0
    for iy in range(ny):
        for iz in range(nz):
            for ix in range(nx):
                if placed >= count:
                    break
                state, a = lcg(state)
                state, b = lcg(state)
                state, c = lcg(state)
                state, d = lcg(state)
                ax, ay, az = a * 2 - 1, b * 2 - 1, c * 2 - 1
                if abs(ax) + abs(ay) + abs(az) < 1e-3:
                    ax = 1.0
                x = (ix - (nx - 1) / 2.0) * spacing
                y = 1.2 + iy * spacing
                z = (iz - (nz - 1) / 2.0) * spacing
                lines.append("push rotate %.6f %.6f %.6f %.3f translate %.4f %.4f %.4f model torus %d %d 0.75 0.25 %s pop"
                             % (ax, ay, az, d * 360.0, x, y, z, tess[0], tess[1], palette[placed % len(palette)][0]))
                placed += 1
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return placed, (nx, ny, nz)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--count", type=int, default=10000)
    ap.add_argument("--tess", type=int, nargs=2, default=(250, 100))
    a = ap.parse_args()
    n, dims = write_scene(a.out, a.count, tuple(a.tess))
    print("wrote %s: %d instances on a %dx%dx%d lattice" % (a.out, n, dims[0], dims[1], dims[2]))
