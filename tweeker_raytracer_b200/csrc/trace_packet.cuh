// trace_packet.cuh -- packet traversal for COHERENT rays: one warp, 32 rays, ONE shared traversal.
//
// Replaces optixTrace for the primary rays (apps/rtigo3/shaders/raygeneration.cu:84-89, first segment).  The path ids of a
// launch are laid out so that a warp covers an 8x4 pixel tile (kernels_shade.cu launch_xy): its 32 primary rays leave the same
// point in nearly the same direction and visit nearly the same nodes.  The one-ray-per-lane driver (trace.cuh trace_stream)
// nevertheless keeps a stack per lane, runs the node test at 26/32 live lanes and the triangle test / instance entry at 3-10
// because the lanes drift apart by a step or two.  Here the warp keeps ONE stack (shared memory, 512 B) and ONE control flow:
//   - a wide node is visited when ANY lane's ray hits it: every lane tests its own ray against the node's eight boxes with its
//     own [tmin, tlimit] (node_test of trace.cuh, 32/32 lanes), the eight hit bits are OR-reduced with one REDUX;
//   - every lane tests every triangle of a visited leaf and keeps its own closest hit; every lane enters every instance of a
//     visited instance-level leaf (the SIMD cost of the transform is the same for 1 or 32 lanes);
//   - node order comes from the direction octant of the first live lane.
// All branches are warp-uniform, node / triangle / instance fetches are warp-uniform addresses (one L1 transaction), and the
// per-lane stack bookkeeping of trace_stream disappears.  Results are IDENTICAL to the per-ray traversal: the closest hit
// (smallest t, ties -> smaller (instance, primitive)) does not depend on the order in which candidates are found, the box tests
// are conservative per ray, and a ray only ever tests MORE triangles than it would alone (those of leaves its neighbours hit).
// The price is the union: a packet visits every node any of its rays needs.  Incoherent rays must use trace_stream.
//
// MEASURED (round 2, profiles/sweeps_r2.md): bit-exact, but slower than the per-ray driver even on primary rays -- geometry
// scene 2359 vs 2496 Msamples/s, Cornell box 742 vs 805, instanced stress scene 237 vs 360 -- because every visited node and
// triangle costs a full-warp test whether one ray or all of them need it, and the 123 registers leave 16 warps per SM for a
// loop that is one long dependent chain (uniform load -> test -> REDUX -> next).  Off by default (RTC_PRIMARY_PACKETS=1).
#pragma once

#include "trace.cuh"

#define RTC_PACKET_STACK 64
#ifndef RTC_PACKETS_PER_FETCH
#define RTC_PACKETS_PER_FETCH 4      // packets a warp takes from the cursor with one atomicAdd
#endif

// every lane pushes the same (warp-uniform) entry; an overflow is counted like the per-ray stack's (rtc_stats::stackOverflows)
__device__ __forceinline__ void packet_push(uint2* __restrict__ stack, uint32_t& sp, const uint2 v, const uint32_t lane)
{
  if (sp < RTC_PACKET_STACK) stack[sp++] = v;
  else if (lane == 0) atomicAdd(&RTC_STACK_OVERFLOW_COUNTER, 1u);
}

// Policy: as trace_stream (load(i, org, dir, tag) -> bool, store(tag, hit)).  smStackWarp: RTC_PACKET_STACK uint2 of this warp.
template <class Policy>
__device__ __forceinline__ void trace_packets(const SceneDesc& sc, uint32_t n, uint32_t* __restrict__ cursor, const Policy& policy,
                                              uint2* __restrict__ smStackWarp)
{
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t fetched = RTC_PACKETS_PER_FETCH, chunk = 0;; ++fetched)
  {
    if (fetched == RTC_PACKETS_PER_FETCH)
    {
      if (lane == 0) chunk = atomicAdd(cursor, 32u * RTC_PACKETS_PER_FETCH);
      chunk = __shfl_sync(0xffffffffu, chunk, 0);
      fetched = 0;
    }
    const uint32_t base = chunk + 32u * fetched;
    if (base >= n) break;
    const uint32_t index = base + lane;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f), d = make_float4(0.f, 0.f, 1.f, -1.f);
    uint32_t tag = 0;
    bool alive = index < n && policy.load(index, o, d, tag);
    const float tmin = o.w;
    float tlimit = d.w;
    float hitT = -1.0f, bV = 0.0f, bW = 0.0f, bDet = 1.0f;
    uint32_t hitInst = 0xffffffffu, hitPrim = 0xffffffffu;
    const bool traced = alive && (tlimit > tmin);      // an empty interval is a miss, stored below
    const uint32_t live = __ballot_sync(0xffffffffu, traced);
    if (live)
    {
      const uint32_t first = (uint32_t)__ffs((int)live) - 1u;
      BoxRay wbr, br;
      box_setup<false>(wbr, o.x, o.y, o.z, d.x, d.y, d.z);
      br = wbr;
      ObjRay orr;
      orr.Sx = orr.Sy = orr.Sz = 0.0f; orr.kx = 0; orr.ky = 1; orr.kz = 2; orr.dx = orr.dy = orr.dz = 0.0f;
      const uint32_t octWorld = __shfl_sync(0xffffffffu, wbr.octinv, first);
      uint32_t oct = octWorld;
      const uint4* nodes = sc.tlasNodes;
      const float4* tris = nullptr;
      uint2 ng = make_uint2(0u, 0x80000000u);
      uint32_t sp = 0u, curInst = 0u;
      int blasBase = -1;
      for (;;)
      {
        uint2 tg;
        if (ng.y & 0xff000000u)
        {
          const uint32_t bit = 31u - (uint32_t)__clz((int)ng.y);
          ng.y &= ~(1u << bit);
          if (ng.y & 0xff000000u) packet_push(smStackWarp, sp, ng, lane);
          const uint32_t slot = (bit - 24u) ^ oct;
          const uint32_t rel = (uint32_t)__popc(ng.y & 0xffu & ((1u << slot) - 1u));
          const uint4* np = nodes + (size_t)(ng.x + rel) * 5u;
          const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
          const uint32_t mine = traced ? node_test(br, n0, n2, n3, n4, tmin, tlimit, n1.x >> 31) : 0u;
          const uint32_t hits = __reduce_or_sync(0xffffffffu, mine);
          const uint32_t imask = n0.w >> 24;
          ng = make_uint2(n1.x, (xor_permute8(hits & imask, oct) << 24) | imask);
          uint32_t leaf = hits & ~imask & 0xffu, primMask = 0u;
          while (leaf)
          {
            const uint32_t s = (uint32_t)__ffs((int)leaf) - 1u;
            leaf &= leaf - 1u;
            const uint32_t meta = (((s < 4u) ? n1.z : n1.w) >> (8u * (s & 3u))) & 0xffu;
            primMask |= ((1u << (meta >> 5)) - 1u) << (meta & 31u);
          }
          tg = make_uint2(n1.y, primMask);
        }
        else
        {
          tg = ng;
          ng = make_uint2(0u, 0u);
        }

        while (tg.y)
        {
          const uint32_t idx = (uint32_t)__ffs((int)tg.y) - 1u;
          tg.y &= tg.y - 1u;
          if (blasBase < 0)
          {
            // instance-level leaf: every live lane enters the instance
            const uint32_t inst = __ldg(sc.tlasLeaves + tg.x + idx);
            if (tg.y) packet_push(smStackWarp, sp, tg, lane);
            if (ng.y & 0xff000000u) packet_push(smStackWarp, sp, ng, lane);
            const float4* ip = sc.instances + (size_t)inst * 4u;
            const float4 r0 = __ldg(ip), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2), r3 = __ldg(ip + 3);
            const float oox = __fmaf_rn(r0.x, o.x, __fmaf_rn(r0.y, o.y, __fmaf_rn(r0.z, o.z, r0.w)));
            const float ooy = __fmaf_rn(r1.x, o.x, __fmaf_rn(r1.y, o.y, __fmaf_rn(r1.z, o.z, r1.w)));
            const float ooz = __fmaf_rn(r2.x, o.x, __fmaf_rn(r2.y, o.y, __fmaf_rn(r2.z, o.z, r2.w)));
            orr.dx = __fmaf_rn(r0.x, d.x, __fmaf_rn(r0.y, d.y, __fmul_rn(r0.z, d.z)));
            orr.dy = __fmaf_rn(r1.x, d.x, __fmaf_rn(r1.y, d.y, __fmul_rn(r1.z, d.z)));
            orr.dz = __fmaf_rn(r2.x, d.x, __fmaf_rn(r2.y, d.y, __fmul_rn(r2.z, d.z)));
            shear_setup(orr);
            box_setup<false>(br, oox, ooy, ooz, orr.dx, orr.dy, orr.dz);
            oct = __shfl_sync(0xffffffffu, br.octinv, first);
            curInst = inst;
            blasBase = (int)sp;
            nodes = reinterpret_cast<const uint4*>(((unsigned long long)__float_as_uint(r3.y) << 32) | __float_as_uint(r3.x));
            tris = reinterpret_cast<const float4*>(((unsigned long long)__float_as_uint(r3.w) << 32) | __float_as_uint(r3.z));
            ng = make_uint2(0u, 0x80000000u);
            tg = make_uint2(0u, 0u);
            break;
          }
          else
          {
            const float4* tp = tris + (size_t)(tg.x + idx) * 3u;
            const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
            float t, det, V, W;
            if (traced && tri_test(orr, br.ox, br.oy, br.oz, v0, v1, v2, t, det, V, W) && t > tmin)
            {
              const uint32_t prim = __float_as_uint(v0.w);
              const bool better = (hitInst != 0xffffffffu) ? (t < hitT || (t == hitT && (curInst < hitInst || (curInst == hitInst && prim < hitPrim))))
                                                           : (t < tlimit);
              if (better)
              {
                tlimit = t; hitT = t; hitInst = curInst; hitPrim = prim;
                bV = V; bW = W; bDet = det;
              }
            }
          }
        }

        if (!(ng.y & 0xff000000u))
        {
          if (blasBase >= 0 && sp == (uint32_t)blasBase)
          {
            blasBase = -1;       // the instance is done: back to the world-space rays
            nodes = sc.tlasNodes;
            br = wbr;
            oct = octWorld;
          }
          if (sp == 0u) break;
          ng = smStackWarp[--sp];
        }
      }
    }
    if (alive)
    {
      TraceHit h;
      h.t = -1.0f; h.u = 0.0f; h.v = 0.0f; h.inst = hitInst; h.prim = 0xffffffffu;
      if (hitInst != 0xffffffffu) { h.t = hitT; h.prim = hitPrim; h.u = __fdiv_rn(bV, bDet); h.v = __fdiv_rn(bW, bDet); }
      policy.store(tag, h);
    }
    __syncwarp();      // the shared stack is reused by the next packet
  }
}
