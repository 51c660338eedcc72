// trace_pool.cuh -- traversal driver with a per-warp RAY POOL: rays are decoupled from lanes.
//
// Replaces optixTrace (apps/rtigo3/shaders/raygeneration.cu:84-89 radiance rays, closesthit.cu:281-286 shadow rays);
// the arithmetic (box test, watertight triangle test, object-space ray, hit ordering) is trace.cuh's, unchanged.
//
// Why: with one ray per lane (trace.cuh trace_stream) a warp runs the node test at 26/32 live lanes but the triangle test
// at 3.3/32 and the instance entry at 10.6/32 (ncu source view, profiles/ncu_trace_r1_final.csv): the lanes of a warp are
// spread over the three phases and every phase is executed for the few lanes that are in it.  Here every warp owns
// R = 32 * RTC_POOL_K ray slots whose complete traversal state lives in shared memory (word-major: word w of slot s at
// sm[w * R + s], so 32 distinct slots are at most K-way bank conflicted).  Each iteration the warp counts the slots per
// phase (one REDUX), picks the fullest phase, compacts up to 32 slots of that phase onto its lanes (ballots + a 32-entry
// list) and runs ONLY that phase:
//   N  visit one wide node        (slot has an inner-node group)
//   T  test the slot's triangles  (slot has a leaf group inside a GAS)
//   I  enter one instance         (slot has a leaf group at the instance level)
//   F  store the result and fetch a new ray into the slot (finished or empty slots, >= RTC_POOL_FETCH of them)
// A ray that changes phase simply waits in its slot while the lanes work on other rays, so nothing is postponed inside a
// ray (same node order, same tlimit updates, same results as the one-ray-per-lane driver).  No barrier wider than a warp.
#pragma once

#include "trace.cuh"

#ifndef RTC_POOL_K
#define RTC_POOL_K 2              // ray slots per lane
#endif
#ifndef RTC_POOL_STACK
#define RTC_POOL_STACK 4          // traversal stack entries per slot kept in shared memory (the rest spills to a global scratch)
#endif
#ifndef RTC_POOL_FETCH
#define RTC_POOL_FETCH 16         // refill when at least this many slots are free (finished or empty)
#endif
#ifndef RTC_POOL_TWEIGHT
#define RTC_POOL_TWEIGHT 2        // a slot in the triangle phase counts as this many lanes of work when the phases are compared
#endif
#define RTC_POOL_TOTAL_STACK 40   // stack entries per ray (shared + global), as in trace.cuh

namespace rtpool {

constexpr uint32_t K = RTC_POOL_K, R = 32u * RTC_POOL_K, STK = RTC_POOL_STACK, OVF = RTC_POOL_TOTAL_STACK - RTC_POOL_STACK;

enum : uint32_t { PH_EMPTY = 0, PH_N = 1, PH_T = 2, PH_I = 3, PH_F = 4 };
enum : int { PASS_N = 0, PASS_T = 1, PASS_I = 2, PASS_F = 3 };

// state words of one slot.  The first six words of each space are laid out alike so that the node pass addresses
// "origin and 1/d of the current space" with one offset.
enum : uint32_t
{
  W_WORG = 0,        // 0-2  world origin
  W_WINV = 3,        // 3-5  world 1/d
  W_OORG = 6,        // 6-8  object origin (inside an instance)
  W_OINV = 9,        // 9-11 object 1/d
  W_WDIR = 12,       // 12-14 world direction
  W_SHEAR = 15,      // 15-17 Sx, Sy, Sz
  W_K = 18,          // kx | ky << 2 | kz << 4 | octinv(object) << 8 | octinv(world) << 12
  W_TMIN = 19, W_TLIMIT = 20,
  W_HITINST = 21, W_HITPRIM = 22,
  W_NGX = 23, W_NGY = 24, W_TGX = 25, W_TGY = 26,
  W_SP = 27,         // sp | (blasBase + 1) << 8 ; bits 8.. zero: instance level
  W_CURINST = 28,
  W_NODES = 29,      // 29-30 node array of the current level
  W_TRIS = 31,       // 31-32 triangle array of the current GAS
  W_TAG = 33,        // the policy's per-ray word (path id)
  W_PHASE = 34,
  W_BARY = 35,       // 35-37 V, W, det of the best hit (closest-hit kernels only)
  NUM_COMMON = 35
};

// words per slot of one kernel flavour: any-hit kernels carry no barycentrics, only SKIP kernels carry the skip key
__host__ __device__ constexpr uint32_t num_words(bool any, bool skip) { return NUM_COMMON + (any ? 0u : 3u) + (skip ? 3u : 0u); }
__host__ __device__ constexpr uint32_t skip_word(bool any) { return NUM_COMMON + (any ? 0u : 3u); }
__host__ __device__ constexpr uint32_t warp_words(bool any, bool skip) { return (num_words(any, skip) + 2u * STK) * R + 32u; }   // state, stack x / y columns, compaction list
__host__ __device__ constexpr uint32_t warp_bytes(bool any, bool skip) { return warp_words(any, skip) * 4u; }
constexpr uint32_t kOverflowPerWarp = R * OVF;                          // uint2 entries of global scratch per warp

template <uint32_t NUM_WORDS>
struct Pool
{
  uint32_t* sm;          // this warp's shared words
  uint2*    ovf;         // this warp's global stack scratch (R * OVF entries)
  __device__ __forceinline__ uint32_t& w(uint32_t word, uint32_t slot) const { return sm[word * R + slot]; }
  __device__ __forceinline__ float& f(uint32_t word, uint32_t slot) const { return reinterpret_cast<float*>(sm)[word * R + slot]; }
  __device__ __forceinline__ uint32_t* list() const { return sm + (NUM_WORDS + 2u * STK) * R; }

  __device__ __forceinline__ void push(uint32_t slot, uint32_t& sp, const uint2 v) const
  {
    if (sp < STK) { sm[(NUM_WORDS + sp) * R + slot] = v.x; sm[(NUM_WORDS + STK + sp) * R + slot] = v.y; }
    else if (sp < STK + OVF) __stcg(ovf + slot * OVF + (sp - STK), v);
    else { atomicAdd(&RTC_STACK_OVERFLOW_COUNTER, 1u); return; }
    ++sp;
  }
  __device__ __forceinline__ uint2 pop(uint32_t slot, uint32_t& sp) const
  {
    --sp;
    if (sp < STK) return make_uint2(sm[(NUM_WORDS + sp) * R + slot], sm[(NUM_WORDS + STK + sp) * R + slot]);
    return __ldcg(ovf + slot * OVF + (sp - STK));
  }
};

// Up to 32 slots whose predicate holds, one per lane: returns how many, slot valid for lane < count.
template <class Pool>
__device__ __forceinline__ uint32_t compact(const Pool& p, const bool (&pred)[K], uint32_t lane, uint32_t& slot)
{
  uint32_t total = 0;
  uint32_t* list = p.list();
#pragma unroll
  for (uint32_t k = 0; k < K; ++k)
  {
    const uint32_t b = __ballot_sync(0xffffffffu, pred[k]);
    const uint32_t pos = total + (uint32_t)__popc(b & ((1u << lane) - 1u));
    if (pred[k] && pos < 32u) list[pos] = lane + 32u * k;
    total += (uint32_t)__popc(b);
  }
  __syncwarp();
  const uint32_t count = total < 32u ? total : 32u;
  slot = lane < count ? list[lane] : 0u;
  return count;
}

// The slot's node group has no inner children left and its leaf group is empty: leave the instance if its subtree is done,
// pop the next group, or finish the ray.  Writes the groups, the stack word and the phase.
template <class Pool>
__device__ __forceinline__ void advance(const Pool& p, const SceneDesc& sc, uint32_t slot, uint2 ng, uint32_t spw)
{
  uint32_t phase;
  if (ng.y & 0xff000000u) { phase = PH_N; p.w(W_SP, slot) = spw; }      // the caller may have pushed
  else
  {
    uint32_t sp = spw & 0xffu, bb = spw >> 8;
    if (bb != 0u && sp == bb - 1u)
    {
      bb = 0u;        // back to the world-space ray
      const unsigned long long tn = (unsigned long long)sc.tlasNodes;
      p.w(W_NODES, slot) = (uint32_t)tn; p.w(W_NODES + 1, slot) = (uint32_t)(tn >> 32);
    }
    if (sp == 0u) { phase = PH_F; }
    else
    {
      const uint2 g = p.pop(slot, sp);
      if (g.y & 0xff000000u) { ng = g; phase = PH_N; }
      else
      {
        p.w(W_TGX, slot) = g.x; p.w(W_TGY, slot) = g.y;
        ng = make_uint2(0u, 0u);
        phase = bb ? PH_T : PH_I;
      }
    }
    p.w(W_SP, slot) = sp | (bb << 8);
  }
  p.w(W_NGX, slot) = ng.x; p.w(W_NGY, slot) = ng.y;
  p.w(W_PHASE, slot) = phase;
}

// Policy: bool load(i, org, dir, tag) (false: skip this index), void store(tag, hit); SKIP kernels also skip_key(tag, t, inst, prim).
template <bool ANY, bool COUNT, bool SKIP, class Policy>
__device__ __forceinline__ void trace_pool(const SceneDesc& sc, uint32_t n, uint32_t* __restrict__ cursor, const Policy& policy,
                                           uint32_t* __restrict__ smWarp, uint2* __restrict__ ovfWarp, unsigned long long* __restrict__ countsOut)
{
  constexpr uint32_t W_SKIP = skip_word(ANY);
  const Pool<num_words(ANY, SKIP)> p = { smWarp, ovfWarp };
  const uint32_t lane = threadIdx.x & 31u;
  bool exhausted = false;
  unsigned long long cNodes = 0, cTris = 0, cInsts = 0, cRays = 0;
  uint32_t cPasses[4] = { 0u, 0u, 0u, 0u }, cLanes[4] = { 0u, 0u, 0u, 0u };      // COUNT: passes per phase and the lanes they occupied (lane 0)
#pragma unroll
  for (uint32_t k = 0; k < K; ++k) p.w(W_PHASE, lane + 32u * k) = PH_EMPTY;

  for (;;)
  {
    __syncwarp();
    // ---- census: slots per phase
    uint32_t ph[K], contrib = 0u;
#pragma unroll
    for (uint32_t k = 0; k < K; ++k)
    {
      ph[k] = p.w(W_PHASE, lane + 32u * k);
      contrib += ph[k] ? (1u << ((ph[k] - 1u) << 3)) : 0u;
    }
    const uint32_t packed = __reduce_add_sync(0xffffffffu, contrib);
    const uint32_t nN = packed & 0xffu, nT = (packed >> 8) & 0xffu, nI = (packed >> 16) & 0xffu, nF = packed >> 24;
    const uint32_t nE = R - nN - nT - nI - nF;
    int pass;
    if (!exhausted && nE + nF >= RTC_POOL_FETCH) pass = PASS_F;
    else
    {
      const uint32_t sN = nN < 32u ? nN : 32u, sT = nT * RTC_POOL_TWEIGHT < 32u ? nT * RTC_POOL_TWEIGHT : 32u, sI = nI < 32u ? nI : 32u;
      if ((sN | sT | sI) == 0u)
      {
        if (nF != 0u || (!exhausted && nE != 0u)) pass = PASS_F;
        else break;
      }
      else if (sN >= sT && sN >= sI) pass = PASS_N;
      else if (sT >= sI) pass = PASS_T;
      else pass = PASS_I;
    }

    bool pred[K];
    uint32_t slot;
    if (pass == PASS_N)
    {
      // ---- visit one wide node per slot
#pragma unroll
      for (uint32_t k = 0; k < K; ++k) pred[k] = ph[k] == PH_N;
      const uint32_t count = compact(p, pred, lane, slot);
      if (COUNT) { cPasses[0]++; cLanes[0] += count; }
      if (lane < count)
      {
        uint2 ng = make_uint2(p.w(W_NGX, slot), p.w(W_NGY, slot));
        uint32_t spw = p.w(W_SP, slot);
        const uint32_t kw = p.w(W_K, slot);
        const bool inGas = (spw >> 8) != 0u;
        const uint32_t space = inGas ? (W_OORG * R) : (W_WORG * R);
        BoxRay br;
        br.ox = reinterpret_cast<const float*>(p.sm)[space + 0u * R + slot];
        br.oy = reinterpret_cast<const float*>(p.sm)[space + 1u * R + slot];
        br.oz = reinterpret_cast<const float*>(p.sm)[space + 2u * R + slot];
        br.idx = reinterpret_cast<const float*>(p.sm)[space + 3u * R + slot];
        br.idy = reinterpret_cast<const float*>(p.sm)[space + 4u * R + slot];
        br.idz = reinterpret_cast<const float*>(p.sm)[space + 5u * R + slot];
        br.octinv = (inGas ? (kw >> 8) : (kw >> 12)) & 7u;
        const float tmin = p.f(W_TMIN, slot), tlimit = p.f(W_TLIMIT, slot);
        const uint4* nodes = reinterpret_cast<const uint4*>(((unsigned long long)p.w(W_NODES + 1, slot) << 32) | p.w(W_NODES, slot));

        const uint32_t bit = 31u - (uint32_t)__clz((int)ng.y);
        ng.y &= ~(1u << bit);
        if (ng.y & 0xff000000u)
        {
          uint32_t sp = spw & 0xffu;
          p.push(slot, sp, ng);
          spw = (spw & 0xffffff00u) | sp;
        }
        const uint32_t cslot = (bit - 24u) ^ br.octinv;
        const uint32_t rel = (uint32_t)__popc(ng.y & 0xffu & ((1u << cslot) - 1u));
        const uint4* np = nodes + (size_t)(ng.x + rel) * 5u;
        const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
        if (COUNT) cNodes++;
        const uint32_t hits = node_test(br, n0, n2, n3, n4, tmin, tlimit, n1.x >> 31);
        const uint32_t imask = n0.w >> 24;
        ng = make_uint2(n1.x, (xor_permute8(hits & imask, br.octinv) << 24) | imask);
        uint32_t leaf = hits & ~imask, primMask = 0u;
        while (leaf)
        {
          const uint32_t s = (uint32_t)__ffs((int)leaf) - 1u;
          leaf &= leaf - 1u;
          const uint32_t meta = (((s < 4u) ? n1.z : n1.w) >> (8u * (s & 3u))) & 0xffu;
          primMask |= ((1u << (meta >> 5)) - 1u) << (meta & 31u);
        }
        if (primMask)
        {
          p.w(W_NGX, slot) = ng.x; p.w(W_NGY, slot) = ng.y;
          p.w(W_TGX, slot) = n1.y; p.w(W_TGY, slot) = primMask;
          p.w(W_SP, slot) = spw;
          p.w(W_PHASE, slot) = inGas ? PH_T : PH_I;
        }
        else advance(p, sc, slot, ng, spw);
      }
    }
    else if (pass == PASS_T)
    {
      // ---- test the triangles of the slot's leaf group
#pragma unroll
      for (uint32_t k = 0; k < K; ++k) pred[k] = ph[k] == PH_T;
      const uint32_t count = compact(p, pred, lane, slot);
      if (COUNT) { cPasses[1]++; cLanes[1] += count; }
      if (lane < count)
      {
        uint2 tg = make_uint2(p.w(W_TGX, slot), p.w(W_TGY, slot));
        const uint32_t kw = p.w(W_K, slot);
        ObjRay orr;
        orr.kx = (int)(kw & 3u); orr.ky = (int)((kw >> 2) & 3u); orr.kz = (int)((kw >> 4) & 3u);
        orr.Sx = p.f(W_SHEAR, slot); orr.Sy = p.f(W_SHEAR + 1, slot); orr.Sz = p.f(W_SHEAR + 2, slot);
        const float ox = p.f(W_OORG, slot), oy = p.f(W_OORG + 1, slot), oz = p.f(W_OORG + 2, slot);
        const float tmin = p.f(W_TMIN, slot);
        float tlimit = p.f(W_TLIMIT, slot);
        uint32_t hitInst = p.w(W_HITINST, slot), hitPrim = p.w(W_HITPRIM, slot);
        const uint32_t curInst = p.w(W_CURINST, slot);
        const float4* tris = reinterpret_cast<const float4*>(((unsigned long long)p.w(W_TRIS + 1, slot) << 32) | p.w(W_TRIS, slot));
        float skipT = 0.0f; uint32_t skipInst = 0u, skipPrim = 0u;
        if (SKIP) { skipT = p.f(W_SKIP, slot); skipInst = p.w(W_SKIP + 1, slot); skipPrim = p.w(W_SKIP + 2, slot); }
        float hitT = tlimit, bV = 0.0f, bW = 0.0f, bDet = 0.0f;      // closest: hitT == tlimit once something was found
        bool changed = false;
        while (tg.y)
        {
          const uint32_t idx = (uint32_t)__ffs((int)tg.y) - 1u;
          tg.y &= tg.y - 1u;
          const float4* tp = tris + (size_t)(tg.x + idx) * 3u;
          const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1), v2 = __ldg(tp + 2);
          if (COUNT) cTris++;
          float t, det, V, W;
          if (tri_test(orr, ox, oy, oz, v0, v1, v2, t, det, V, W) && t > tmin)
          {
            const uint32_t prim = __float_as_uint(v0.w);
            if (SKIP && !(t > skipT || (t == skipT && (curInst > skipInst || (curInst == skipInst && prim > skipPrim))))) continue;
            if (ANY)
            {
              if (t < tlimit) { hitT = t; hitInst = curInst; hitPrim = prim; changed = true; break; }
            }
            else
            {
              const bool better = (hitInst != 0xffffffffu) ? (t < hitT || (t == hitT && (curInst < hitInst || (curInst == hitInst && prim < hitPrim))))
                                                           : (t < tlimit);
              if (better)
              {
                tlimit = t; hitT = t; hitInst = curInst; hitPrim = prim;
                bV = V; bW = W; bDet = det;
                changed = true;
              }
            }
          }
        }
        if (changed)
        {
          p.f(W_TLIMIT, slot) = hitT;
          p.w(W_HITINST, slot) = hitInst; p.w(W_HITPRIM, slot) = hitPrim;
          if (!ANY) { p.f(W_BARY, slot) = bV; p.f(W_BARY + 1, slot) = bW; p.f(W_BARY + 2, slot) = bDet; }
        }
        if (ANY && changed) p.w(W_PHASE, slot) = PH_F;       // first hit ends a shadow ray
        else advance(p, sc, slot, make_uint2(p.w(W_NGX, slot), p.w(W_NGY, slot)), p.w(W_SP, slot));
      }
    }
    else if (pass == PASS_I)
    {
      // ---- enter one instance of the slot's instance-level leaf group
#pragma unroll
      for (uint32_t k = 0; k < K; ++k) pred[k] = ph[k] == PH_I;
      const uint32_t count = compact(p, pred, lane, slot);
      if (COUNT) { cPasses[2]++; cLanes[2] += count; }
      if (lane < count)
      {
        uint2 tg = make_uint2(p.w(W_TGX, slot), p.w(W_TGY, slot));
        const uint2 ng = make_uint2(p.w(W_NGX, slot), p.w(W_NGY, slot));
        uint32_t sp = p.w(W_SP, slot) & 0xffu;
        const uint32_t idx = (uint32_t)__ffs((int)tg.y) - 1u;
        tg.y &= tg.y - 1u;
        const uint32_t inst = __ldg(sc.tlasLeaves + tg.x + idx);
        if (tg.y) p.push(slot, sp, tg);
        if (ng.y & 0xff000000u) p.push(slot, sp, ng);
        const float4* ip = sc.instances + (size_t)inst * 4u;
        const float4 r0 = __ldg(ip), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2), r3 = __ldg(ip + 3);
        if (COUNT) cInsts++;
        const float wox = p.f(W_WORG, slot), woy = p.f(W_WORG + 1, slot), woz = p.f(W_WORG + 2, slot);
        const float wdx = p.f(W_WDIR, slot), wdy = p.f(W_WDIR + 1, slot), wdz = p.f(W_WDIR + 2, slot);
        const float oox = __fmaf_rn(r0.x, wox, __fmaf_rn(r0.y, woy, __fmaf_rn(r0.z, woz, r0.w)));
        const float ooy = __fmaf_rn(r1.x, wox, __fmaf_rn(r1.y, woy, __fmaf_rn(r1.z, woz, r1.w)));
        const float ooz = __fmaf_rn(r2.x, wox, __fmaf_rn(r2.y, woy, __fmaf_rn(r2.z, woz, r2.w)));
        ObjRay orr;
        orr.dx = __fmaf_rn(r0.x, wdx, __fmaf_rn(r0.y, wdy, __fmul_rn(r0.z, wdz)));
        orr.dy = __fmaf_rn(r1.x, wdx, __fmaf_rn(r1.y, wdy, __fmul_rn(r1.z, wdz)));
        orr.dz = __fmaf_rn(r2.x, wdx, __fmaf_rn(r2.y, wdy, __fmul_rn(r2.z, wdz)));
        shear_setup(orr);
        BoxRay br;
        box_setup<COUNT>(br, oox, ooy, ooz, orr.dx, orr.dy, orr.dz);
        p.f(W_OORG, slot) = oox; p.f(W_OORG + 1, slot) = ooy; p.f(W_OORG + 2, slot) = ooz;
        p.f(W_OINV, slot) = br.idx; p.f(W_OINV + 1, slot) = br.idy; p.f(W_OINV + 2, slot) = br.idz;
        p.f(W_SHEAR, slot) = orr.Sx; p.f(W_SHEAR + 1, slot) = orr.Sy; p.f(W_SHEAR + 2, slot) = orr.Sz;
        const uint32_t kw = p.w(W_K, slot);
        p.w(W_K, slot) = (kw & 0xf000u) | (uint32_t)orr.kx | ((uint32_t)orr.ky << 2) | ((uint32_t)orr.kz << 4) | (br.octinv << 8);
        p.w(W_CURINST, slot) = inst;
        p.w(W_SP, slot) = sp | ((sp + 1u) << 8);
        p.w(W_NODES, slot) = __float_as_uint(r3.x); p.w(W_NODES + 1, slot) = __float_as_uint(r3.y);
        p.w(W_TRIS, slot) = __float_as_uint(r3.z); p.w(W_TRIS + 1, slot) = __float_as_uint(r3.w);
        p.w(W_NGX, slot) = 0u; p.w(W_NGY, slot) = 0x80000000u;
        p.w(W_TGX, slot) = 0u; p.w(W_TGY, slot) = 0u;
        p.w(W_PHASE, slot) = PH_N;
      }
    }
    else
    {
      // ---- store finished rays, fetch new ones into the freed slots
#pragma unroll
      for (uint32_t k = 0; k < K; ++k) pred[k] = ph[k] == PH_F || (ph[k] == PH_EMPTY && !exhausted);
      const uint32_t count = compact(p, pred, lane, slot);
      if (COUNT) { cPasses[3]++; cLanes[3] += count; }
      if (lane < count && p.w(W_PHASE, slot) == PH_F)
      {
        TraceHit h;
        h.t = -1.0f; h.u = 0.0f; h.v = 0.0f; h.inst = p.w(W_HITINST, slot); h.prim = 0xffffffffu;
        if (h.inst != 0xffffffffu)
        {
          h.t = p.f(W_TLIMIT, slot); h.prim = p.w(W_HITPRIM, slot);
          if (!ANY) { const float det = p.f(W_BARY + 2, slot); h.u = __fdiv_rn(p.f(W_BARY, slot), det); h.v = __fdiv_rn(p.f(W_BARY + 1, slot), det); }
        }
        policy.store(p.w(W_TAG, slot), h);
        if (COUNT) cRays++;
      }
      uint32_t newPhase = PH_EMPTY;
      if (!exhausted)
      {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(cursor, count);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base + count >= n) exhausted = true;
        const uint32_t index = base + lane;
        if (lane < count && index < n)
        {
          float4 o, d; uint32_t tag;
          if (policy.load(index, o, d, tag))
          {
            p.w(W_TAG, slot) = tag;
            p.w(W_HITINST, slot) = 0xffffffffu; p.w(W_HITPRIM, slot) = 0xffffffffu;
            p.f(W_TMIN, slot) = o.w; p.f(W_TLIMIT, slot) = d.w;
            if (!(d.w > o.w)) newPhase = PH_F;      // empty interval: a miss
            else
            {
              BoxRay br;
              box_setup<COUNT>(br, o.x, o.y, o.z, d.x, d.y, d.z);
              p.f(W_WORG, slot) = o.x; p.f(W_WORG + 1, slot) = o.y; p.f(W_WORG + 2, slot) = o.z;
              p.f(W_WINV, slot) = br.idx; p.f(W_WINV + 1, slot) = br.idy; p.f(W_WINV + 2, slot) = br.idz;
              p.f(W_WDIR, slot) = d.x; p.f(W_WDIR + 1, slot) = d.y; p.f(W_WDIR + 2, slot) = d.z;
              p.w(W_K, slot) = br.octinv << 12;
              p.w(W_NGX, slot) = 0u; p.w(W_NGY, slot) = 0x80000000u;
              p.w(W_TGX, slot) = 0u; p.w(W_TGY, slot) = 0u;
              p.w(W_SP, slot) = 0u;
              const unsigned long long tn = (unsigned long long)sc.tlasNodes;
              p.w(W_NODES, slot) = (uint32_t)tn; p.w(W_NODES + 1, slot) = (uint32_t)(tn >> 32);
              if constexpr (SKIP)
              {
                float st; uint32_t si, sp2;
                policy.skip_key(tag, st, si, sp2);
                p.f(W_SKIP, slot) = st; p.w(W_SKIP + 1, slot) = si; p.w(W_SKIP + 2, slot) = sp2;
              }
              newPhase = PH_N;
            }
          }
        }
      }
      if (lane < count) p.w(W_PHASE, slot) = newPhase;
    }
  }

  if (COUNT)
  {
    for (int off = 16; off; off >>= 1)
    {
      cNodes += __shfl_down_sync(0xffffffffu, cNodes, off); cTris += __shfl_down_sync(0xffffffffu, cTris, off);
      cInsts += __shfl_down_sync(0xffffffffu, cInsts, off); cRays += __shfl_down_sync(0xffffffffu, cRays, off);
    }
    if (lane == 0 && cRays)
    {
      atomicAdd(countsOut + 0, cNodes); atomicAdd(countsOut + 1, cTris); atomicAdd(countsOut + 2, cInsts); atomicAdd(countsOut + 3, cRays);
#pragma unroll
      for (int k = 0; k < 4; ++k) { atomicAdd(countsOut + 4 + k, (unsigned long long)cPasses[k]); atomicAdd(countsOut + 8 + k, (unsigned long long)cLanes[k]); }
    }
  }
}

} // namespace rtpool
