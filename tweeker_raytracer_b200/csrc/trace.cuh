// trace.cuh -- stack-based traversal of the two-level wide BVH and the watertight ray/triangle test.
//
// Replaces optixTrace (apps/rtigo3/shaders/raygeneration.cu:84-89 radiance rays, closesthit.cu:281-286
// shadow rays) and __anyhit__shadow (anyhit.cu:84-91).  B200 has no RT cores: boxes and triangles are
// tested on the FP32/INT pipes, nodes and triangles are fetched with 128-bit loads.
//
// The ARITHMETIC of the triangle test, of the object-space ray and of the hit ordering is a definition
// shared with the scalar oracle (oracle/rt_oracle.c "The ray/triangle test"); it is written with
// explicit round-to-nearest intrinsics so that the compiler's FMA contraction cannot change it.
// The box test is only required to be conservative and is free to use contracted arithmetic.
#pragma once

#include "rtc_internal.h"

#define RTC_STACK_SIZE 40

struct TraceHit
{
  float    t, u, v;
  uint32_t inst, prim;
};

struct ObjRay
{
  float ox, oy, oz, dx, dy, dz;
  float Sx, Sy, Sz;
  int   kx, ky, kz;
};

__device__ __forceinline__ float sel3(float x, float y, float z, int k) { return k == 0 ? x : (k == 1 ? y : z); }

__device__ __forceinline__ void shear_setup(ObjRay& r)
{
  const float ax = fabsf(r.dx), ay = fabsf(r.dy), az = fabsf(r.dz);
  int kz = (ax >= ay && ax >= az) ? 0 : ((ay >= az) ? 1 : 2);
  int kx = kz + 1; if (kx == 3) kx = 0;
  int ky = kx + 1; if (ky == 3) ky = 0;
  const float dz = sel3(r.dx, r.dy, r.dz, kz);
  if (dz < 0.0f) { const int t = kx; kx = ky; ky = t; }
  r.kx = kx; r.ky = ky; r.kz = kz;
  r.Sx = __fdiv_rn(sel3(r.dx, r.dy, r.dz, kx), dz);
  r.Sy = __fdiv_rn(sel3(r.dx, r.dy, r.dz, ky), dz);
  r.Sz = __fdiv_rn(1.0f, dz);
}

// Woop/Benthin/Wald watertight test; see the oracle for the definition this mirrors operation by operation.
__device__ __forceinline__ bool tri_test(const ObjRay& r, const float4 v0, const float4 v1, const float4 v2,
                                         float& t, float& det, float& V, float& W)
{
  const float A0 = __fsub_rn(v0.x, r.ox), A1 = __fsub_rn(v0.y, r.oy), A2 = __fsub_rn(v0.z, r.oz);
  const float B0 = __fsub_rn(v1.x, r.ox), B1 = __fsub_rn(v1.y, r.oy), B2 = __fsub_rn(v1.z, r.oz);
  const float C0 = __fsub_rn(v2.x, r.ox), C1 = __fsub_rn(v2.y, r.oy), C2 = __fsub_rn(v2.z, r.oz);
  const float Akz = sel3(A0, A1, A2, r.kz), Bkz = sel3(B0, B1, B2, r.kz), Ckz = sel3(C0, C1, C2, r.kz);
  const float Ax = __fmaf_rn(-r.Sx, Akz, sel3(A0, A1, A2, r.kx)), Ay = __fmaf_rn(-r.Sy, Akz, sel3(A0, A1, A2, r.ky));
  const float Bx = __fmaf_rn(-r.Sx, Bkz, sel3(B0, B1, B2, r.kx)), By = __fmaf_rn(-r.Sy, Bkz, sel3(B0, B1, B2, r.ky));
  const float Cx = __fmaf_rn(-r.Sx, Ckz, sel3(C0, C1, C2, r.kx)), Cy = __fmaf_rn(-r.Sy, Ckz, sel3(C0, C1, C2, r.ky));
  float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
  V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
  W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
  if (U == 0.0f || V == 0.0f || W == 0.0f)
  {
    U = (float)__dsub_rn(__dmul_rn((double)Cx, (double)By), __dmul_rn((double)Cy, (double)Bx));
    V = (float)__dsub_rn(__dmul_rn((double)Ax, (double)Cy), __dmul_rn((double)Ay, (double)Cx));
    W = (float)__dsub_rn(__dmul_rn((double)Bx, (double)Ay), __dmul_rn((double)By, (double)Ax));
  }
  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
  det = __fadd_rn(__fadd_rn(U, V), W);
  if (det == 0.0f) return false;
  const float Az = __fmul_rn(r.Sz, Akz), Bz = __fmul_rn(r.Sz, Bkz), Cz = __fmul_rn(r.Sz, Ckz);
  const float T = __fmaf_rn(U, Az, __fmaf_rn(V, Bz, __fmul_rn(W, Cz)));
  t = __fdiv_rn(T, det);
  return true;
}

// Per-space constants of the box test: t = q * adj + org per plane.
struct BoxRay
{
  float idx, idy, idz;       // 1/d with |d| clamped to 2^-80
  float ox, oy, oz;          // origin
  uint32_t octinv;           // 7 ^ octant, octant bit2 = dx<0, bit1 = dy<0, bit0 = dz<0
};

__device__ __forceinline__ float safe_rcp(float d)
{
  if (fabsf(d) < 0x1p-80f) d = copysignf(0x1p-80f, d);
  return __fdiv_rn(1.0f, d);
}

__device__ __forceinline__ void box_setup(BoxRay& b, float ox, float oy, float oz, float dx, float dy, float dz)
{
  b.idx = safe_rcp(dx); b.idy = safe_rcp(dy); b.idz = safe_rcp(dz);
  b.ox = ox; b.oy = oy; b.oz = oz;
  const uint32_t oct = ((dx < 0.0f) ? 4u : 0u) | ((dy < 0.0f) ? 2u : 0u) | ((dz < 0.0f) ? 1u : 0u);
  b.octinv = 7u ^ oct;
}

__device__ __forceinline__ float byte_f(uint32_t w, int i) { return (float)((w >> (8 * i)) & 0xffu); }

// Tests the 8 quantised child boxes of one node; returns the hit mask:
// bits 24..31 inner children at priority (slot ^ octinv), bits 0..23 leaf primitives (relative to triBase).
__device__ __forceinline__ uint32_t node_test(const BoxRay& b, const uint4 n0, const uint4 n1, const uint4 n2, const uint4 n3, const uint4 n4,
                                              float tmin, float tmax)
{
  const uint32_t e = n0.w;
  const float sx = __uint_as_float((e & 0xffu) << 23), sy = __uint_as_float(((e >> 8) & 0xffu) << 23), sz = __uint_as_float(((e >> 16) & 0xffu) << 23);
  const uint32_t imask = e >> 24;
  const float adjx = sx * b.idx, adjy = sy * b.idy, adjz = sz * b.idz;
  const float orgx = (__uint_as_float(n0.x) - b.ox) * b.idx;
  const float orgy = (__uint_as_float(n0.y) - b.oy) * b.idy;
  const float orgz = (__uint_as_float(n0.z) - b.oz) * b.idz;
  // near/far plane words per axis, selected by the direction sign
  const bool nx = b.idx < 0.0f, ny = b.idy < 0.0f, nz = b.idz < 0.0f;
  // layout: n2 = qlox[0..3], qlox[4..7], qloy[0..3], qloy[4..7]; n3 = qloz, qhix; n4 = qhiy, qhiz
  const uint32_t lox0 = n2.x, lox1 = n2.y, loy0 = n2.z, loy1 = n2.w;
  const uint32_t loz0 = n3.x, loz1 = n3.y, hix0 = n3.z, hix1 = n3.w;
  const uint32_t hiy0 = n4.x, hiy1 = n4.y, hiz0 = n4.z, hiz1 = n4.w;
  const uint32_t nearx0 = nx ? hix0 : lox0, nearx1 = nx ? hix1 : lox1, farx0 = nx ? lox0 : hix0, farx1 = nx ? lox1 : hix1;
  const uint32_t neary0 = ny ? hiy0 : loy0, neary1 = ny ? hiy1 : loy1, fary0 = ny ? loy0 : hiy0, fary1 = ny ? loy1 : hiy1;
  const uint32_t nearz0 = nz ? hiz0 : loz0, nearz1 = nz ? hiz1 : loz1, farz0 = nz ? loz0 : hiz0, farz1 = nz ? loz1 : hiz1;
  const uint32_t meta0 = n1.z, meta1 = n1.w;
  const float tmaxPad = tmax * (1.0f + 0x1p-17f);
  uint32_t mask = 0;
#pragma unroll
  for (int s = 0; s < 8; ++s)
  {
    const int i = s & 3;
    const uint32_t wnx = (s < 4) ? nearx0 : nearx1, wfx = (s < 4) ? farx0 : farx1;
    const uint32_t wny = (s < 4) ? neary0 : neary1, wfy = (s < 4) ? fary0 : fary1;
    const uint32_t wnz = (s < 4) ? nearz0 : nearz1, wfz = (s < 4) ? farz0 : farz1;
    const float t0x = fmaf(byte_f(wnx, i), adjx, orgx), t1x = fmaf(byte_f(wfx, i), adjx, orgx);
    const float t0y = fmaf(byte_f(wny, i), adjy, orgy), t1y = fmaf(byte_f(wfy, i), adjy, orgy);
    const float t0z = fmaf(byte_f(wnz, i), adjz, orgz), t1z = fmaf(byte_f(wfz, i), adjz, orgz);
    const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
    const float tf = fminf(fminf(t1x, t1y), t1z) * (1.0f + 0x1p-17f);
    const uint32_t meta = (((s < 4) ? meta0 : meta1) >> (8 * i)) & 0xffu;
    const bool inner = (imask >> s) & 1u;
    if (tn <= fminf(tf, tmaxPad) && (inner || meta != 0u))
    {
      if (inner) mask |= 1u << (24u + ((uint32_t)s ^ b.octinv));
      else       mask |= ((1u << (meta >> 5)) - 1u) << (meta & 31u);
    }
  }
  return mask;
}

// Per-ray work counters of the counting variant (the algorithmic-bytes figure of DESIGN.md section 5).
struct TraceCounts { uint32_t nodes, tris, insts; };

// Closest hit (ANY = false) or first hit (ANY = true) of one ray against the two-level scene.
// Closest hit: smallest t in (tmin, tmax), ties -> smaller (instance, primitive).
template <bool ANY, bool COUNT = false>
__device__ __forceinline__ bool trace_ray(const SceneDesc& sc, const float4 org, const float4 dir, TraceHit& hit,
                                          TraceCounts* counts = nullptr)
{
  const float tmin = org.w;
  float tlimit = dir.w;               // current far bound (shrinks to the best t for closest hit)
  bool found = false;
  hit.t = -1.0f; hit.u = 0.0f; hit.v = 0.0f; hit.inst = 0xffffffffu; hit.prim = 0xffffffffu;
  if (!(tlimit > tmin)) return false;

  uint2 stack[RTC_STACK_SIZE];
  int sp = 0;
  int blasBase = -1;                  // >= 0 while inside an instance: stack height at entry
  uint32_t curInst = 0;
  BoxRay br;
  box_setup(br, org.x, org.y, org.z, dir.x, dir.y, dir.z);
  ObjRay orr;
  const uint4*  nodes = sc.tlasNodes; // node array of the current level
  const float4* tris = nullptr;       // triangle array of the current GAS

  uint2 nodeGroup = make_uint2(0u, 0x80000000u);
  uint2 triGroup = make_uint2(0u, 0u);

  for (;;)
  {
    if (nodeGroup.y & 0xff000000u)
    {
      const uint32_t bit = 31u - (uint32_t)__clz((int)nodeGroup.y);
      nodeGroup.y &= ~(1u << bit);
      if (nodeGroup.y & 0xff000000u) { if (sp < RTC_STACK_SIZE) stack[sp++] = nodeGroup; }
      const uint32_t slot = (bit - 24u) ^ br.octinv;
      const uint32_t rel = (uint32_t)__popc(nodeGroup.y & 0xffu & ((1u << slot) - 1u));
      const uint4* np = nodes + (size_t)(nodeGroup.x + rel) * 5u;
      const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
      if (COUNT) counts->nodes++;
      const uint32_t m = node_test(br, n0, n1, n2, n3, n4, tmin, tlimit);
      nodeGroup = make_uint2(n1.x, (m & 0xff000000u) | (n0.w >> 24));
      triGroup = make_uint2(n1.y, m & 0x00ffffffu);
    }
    else
    {
      triGroup = nodeGroup;
      nodeGroup = make_uint2(0u, 0u);
    }

    while (triGroup.y)
    {
      const uint32_t idx = (uint32_t)__ffs((int)triGroup.y) - 1u;
      triGroup.y &= triGroup.y - 1u;
      if (blasBase < 0)
      {
        // instance-level leaf: enter the instance
        const uint32_t inst = __ldg(sc.tlasLeaves + triGroup.x + idx);
        if (triGroup.y) { if (sp < RTC_STACK_SIZE) stack[sp++] = triGroup; }
        if (nodeGroup.y & 0xff000000u) { if (sp < RTC_STACK_SIZE) stack[sp++] = nodeGroup; }
        const float4* ip = sc.instances + (size_t)inst * 4u;
        const float4 r0 = __ldg(ip), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2), r3 = __ldg(ip + 3);
        if (COUNT) counts->insts++;
        orr.ox = __fmaf_rn(r0.x, org.x, __fmaf_rn(r0.y, org.y, __fmaf_rn(r0.z, org.z, r0.w)));
        orr.oy = __fmaf_rn(r1.x, org.x, __fmaf_rn(r1.y, org.y, __fmaf_rn(r1.z, org.z, r1.w)));
        orr.oz = __fmaf_rn(r2.x, org.x, __fmaf_rn(r2.y, org.y, __fmaf_rn(r2.z, org.z, r2.w)));
        orr.dx = __fmaf_rn(r0.x, dir.x, __fmaf_rn(r0.y, dir.y, __fmul_rn(r0.z, dir.z)));
        orr.dy = __fmaf_rn(r1.x, dir.x, __fmaf_rn(r1.y, dir.y, __fmul_rn(r1.z, dir.z)));
        orr.dz = __fmaf_rn(r2.x, dir.x, __fmaf_rn(r2.y, dir.y, __fmul_rn(r2.z, dir.z)));
        shear_setup(orr);
        box_setup(br, orr.ox, orr.oy, orr.oz, orr.dx, orr.dy, orr.dz);
        curInst = inst;
        blasBase = sp;
        nodes = reinterpret_cast<const uint4*>(((unsigned long long)__float_as_uint(r3.y) << 32) | __float_as_uint(r3.x));
        tris  = reinterpret_cast<const float4*>(((unsigned long long)__float_as_uint(r3.w) << 32) | __float_as_uint(r3.z));
        nodeGroup = make_uint2(0u, 0x80000000u);
        triGroup = make_uint2(0u, 0u);
        break;
      }
      else
      {
        const float4* tp = tris + (size_t)(triGroup.x + idx) * 3u;
        const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
        if (COUNT) counts->tris++;
        float t, det, V, W;
        if (tri_test(orr, v0, v1, v2, t, det, V, W) && t > tmin)
        {
          const uint32_t prim = __float_as_uint(v0.w);
          if (ANY)
          {
            if (t < tlimit) { hit.t = t; hit.inst = curInst; hit.prim = prim; return true; }
          }
          else
          {
            const bool better = found ? (t < hit.t || (t == hit.t && (curInst < hit.inst || (curInst == hit.inst && prim < hit.prim))))
                                      : (t < tlimit);
            if (better)
            {
              found = true; tlimit = t;
              hit.t = t; hit.u = __fdiv_rn(V, det); hit.v = __fdiv_rn(W, det); hit.inst = curInst; hit.prim = prim;
            }
          }
        }
      }
    }

    if (!(nodeGroup.y & 0xff000000u))
    {
      if (blasBase >= 0 && sp == blasBase)
      {
        blasBase = -1;   // leave the instance: back to the world-space ray
        nodes = sc.tlasNodes;
        box_setup(br, org.x, org.y, org.z, dir.x, dir.y, dir.z);
      }
      if (sp == 0) break;
      nodeGroup = stack[--sp];
    }
  }
  return found;
}
