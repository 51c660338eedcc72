// tests/native/trace_host.cpp -- TEST INFRASTRUCTURE: the PRODUCT's traversal source compiled for the host.
//
// csrc/trace.cuh (node test, watertight triangle test, instance entry, the stack machine of Traversal::step, the ordered
// any-hit SKIP variant) is included here unchanged and compiled by g++: the CUDA intrinsics it uses are given their IEEE
// meaning below, its three inline-PTX helpers carry a C++ twin behind `#if defined(__CUDACC__)` (the device build is
// untouched: identical SASS).  One host thread plays one lane: begin(), step() until done, result().  What is NOT covered is
// the warp-level driver (trace_stream: ballots, the ray cursor) -- it is never instantiated here -- and the approximate
// reciprocal of the timed kernels (this build uses the IEEE one, like the GPU's counting kernels).
//
// tests/test_cpu_trace_source.py holds the result against the scalar oracle: hits bit for bit against the oracle's binary
// BVH and brute force, node / triangle / instance counters against oracle/wide_bvh.inc (the restatement the GPU counters are
// compared with), and the SKIP enumeration against orc_trace_closest_after.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>
#include <device_launch_parameters.h>

// ---- the intrinsics trace.cuh uses, with their defined (round-to-nearest) meaning; built with -ffp-contract=off ----------
static inline float  __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float  __fadd_rn(float a, float b) { return a + b; }
static inline float  __fsub_rn(float a, float b) { return a - b; }
static inline float  __fmul_rn(float a, float b) { return a * b; }
static inline float  __fdiv_rn(float a, float b) { return a / b; }
static inline float  __frcp_rn(float a) { return 1.0f / a; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline float    __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned old = *p; *p += v; return old; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long old = *p; *p += v; return old; }
// warp intrinsics: named by trace_stream (a template that is parsed but never instantiated here)
unsigned __ballot_sync(unsigned, int);
unsigned __shfl_sync(unsigned, unsigned, int);
unsigned long long __shfl_down_sync(unsigned, unsigned long long, int);

#include "trace.cuh"

// The exported arrays as the kernels see them: 16-byte aligned copies (nodes, triangles and instance records are read as uint4 /
// float4) and instance records as rtc_ias_build writes them (world->object rows 0..2, {GAS nodes pointer, GAS triangles pointer}).
struct HostScene
{
  SceneDesc sc{};
  std::vector<void*> owned;
  ~HostScene() { for (void* p : owned) free(p); }
  void* aligned(size_t bytes) { void* p = nullptr; if (posix_memalign(&p, 16, bytes ? bytes : 16)) p = nullptr; owned.push_back(p); return p; }
  void* copy(const void* src, size_t bytes) { void* p = aligned(bytes); if (bytes) std::memcpy(p, src, bytes); return p; }
};

extern "C" {

// the layout of orc_wide_scene (oracle/wide_bvh.inc), i.e. of oracle/orc.py WideScene
struct th_scene
{
  const void*     tlasNodes;
  const uint32_t* tlasLeaves;
  const float*    worldToObject;
  const uint32_t* instGas;
  const void* const*  gasNodes;
  const float* const* gasTris;
  uint32_t numInstances;
};

static void build_host_scene(HostScene& hs, const th_scene* ws, uint32_t numTlasNodes, uint32_t numTlasLeaves, uint32_t numGas,
                             const uint32_t* numGasNodes, const uint32_t* numGasTris)
{
  std::vector<const uint4*> gasNodes(numGas);
  std::vector<const float4*> gasTris(numGas);
  for (uint32_t g = 0; g < numGas; ++g)
  {
    gasNodes[g] = (const uint4*)hs.copy(ws->gasNodes[g], (size_t)numGasNodes[g] * 80u);
    gasTris[g] = (const float4*)hs.copy(ws->gasTris[g], (size_t)numGasTris[g] * 48u);
  }
  float4* inst = (float4*)hs.aligned((size_t)ws->numInstances * 64u);
  for (uint32_t i = 0; i < ws->numInstances; ++i)
  {
    const float* m = ws->worldToObject + 12u * (size_t)i;
    for (int r = 0; r < 3; ++r) inst[4u * i + r] = make_float4(m[4 * r], m[4 * r + 1], m[4 * r + 2], m[4 * r + 3]);
    const uint64_t np = (uint64_t)(uintptr_t)gasNodes[ws->instGas[i]], tp = (uint64_t)(uintptr_t)gasTris[ws->instGas[i]];
    const uint32_t w[4] = { (uint32_t)np, (uint32_t)(np >> 32), (uint32_t)tp, (uint32_t)(tp >> 32) };
    std::memcpy(&inst[4u * i + 3], w, 16);
  }
  hs.sc.tlasNodes = (const uint4*)hs.copy(ws->tlasNodes, (size_t)numTlasNodes * 80u);
  hs.sc.tlasLeaves = (const uint32_t*)hs.copy(ws->tlasLeaves, (size_t)numTlasLeaves * 4u);
  hs.sc.instances = inst;
  hs.sc.numInstances = ws->numInstances; hs.sc.numTlasNodes = numTlasNodes; hs.sc.numTlasLeaves = numTlasLeaves;
}

// rays: rtc_ray; hits: rtc_hit; counts: nodes, triangles, instances.  skip: optional, per ray {t bits, instance, primitive}
// -- when given, the SKIP instantiation runs (closest candidate AFTER the key) with tmin raised exactly as the device's
// ExtendPathsAfter::load does.  numTlasNodes / numGasNodes / numGasTris size the 16-byte aligned copies.
int th_trace(const th_scene* ws, uint32_t numTlasNodes, uint32_t numTlasLeaves, uint32_t numGas, const uint32_t* numGasNodes,
             const uint32_t* numGasTris, const rtc_ray* rays, uint64_t n, int any, const uint32_t* skip, rtc_hit* hits, uint64_t counts[3],
             uint64_t* stackOverflows)
{
  HostScene hs;
  build_host_scene(hs, ws, numTlasNodes, numTlasLeaves, numGas, numGasNodes, numGasTris);
  const SceneDesc& sc = hs.sc;

  counts[0] = counts[1] = counts[2] = 0;
  const unsigned overflowsBefore = RTC_STACK_OVERFLOW_COUNTER;
  uint2 smStack[RTC_SM_STACK]; float smRay[RTC_SM_RAY_WORDS]; uint2 lmStack[RTC_LM_STACK];
  auto run = [&](auto& tr, const rtc_ray& r, float tmin) {
    tr.smStack = smStack; tr.smRay = smRay; tr.lmStack = lmStack;
    if (tr.begin(sc, make_float4(r.ox, r.oy, r.oz, tmin), make_float4(r.dx, r.dy, r.dz, r.tmax)))
    {
      while (tr.step(sc)) {}
      counts[0] += tr.counts.nodes; counts[1] += tr.counts.tris; counts[2] += tr.counts.insts;
    }
    return tr.result();
  };
  for (uint64_t i = 0; i < n; ++i)
  {
    TraceHit h;
    if (skip)
    {
      Traversal<false, true, 1, true> tr;
      tr.skipT = __uint_as_float(skip[3 * i]); tr.skipInst = skip[3 * i + 1]; tr.skipPrim = skip[3 * i + 2];
      h = run(tr, rays[i], fmaxf(rays[i].tmin, __uint_as_float(skip[3 * i] - 1u)));
    }
    else if (any) { Traversal<true, true, 1, false> tr; h = run(tr, rays[i], rays[i].tmin); }
    else          { Traversal<false, true, 1, false> tr; h = run(tr, rays[i], rays[i].tmin); }
    if (hits) { hits[i].t = h.t; hits[i].u = h.u; hits[i].v = h.v; hits[i].inst = h.inst; hits[i].prim = h.prim; }
  }
  if (stackOverflows) *stackOverflows = RTC_STACK_OVERFLOW_COUNTER - overflowsBefore;
  return 0;
}

// Lock-step emulation of ONE persistent warp of trace_stream (csrc/trace.cuh) over a ray list, for tests/tools/simd_cost.py: 32
// lanes, each owning one ray; when at least `fetchThreshold` lanes are idle (or all are) the idle lanes take the next rays in
// list order; every iteration every active lane runs Traversal::step (leafThreshold > 0: the gated leaf phase of trace_stream's
// RTC_LEAF_THRESHOLD path instead).  What a warp pays per iteration is the MAXIMUM over its
// lanes, not the sum: out[] accumulates, over all iterations,
//   [0] iterations  [1] iterations in which some lane visited a node  [2] sum over iterations of the largest number of
//   triangles one lane tested  [3] iterations in which some lane entered an instance  [4] lane-steps (active lanes summed)
//   [5] node visits summed over lanes  [6] triangle tests summed over lanes  [7] instance entries summed over lanes
//   [8] refills  [9] rays
// so that warpCost = cN*[1] + cT*[2] + cI*[3] + c0*[0] can be compared between acceleration structures built in different
// ways -- a SIMD-aware version of the per-ray counters, still a model (no memory system, no issue scheduling).
int th_simd_cost(const th_scene* ws, uint32_t numTlasNodes, uint32_t numTlasLeaves, uint32_t numGas, const uint32_t* numGasNodes,
                 const uint32_t* numGasTris, const rtc_ray* rays, uint64_t n, int any, uint32_t fetchThreshold, uint32_t leafThreshold,
                 uint64_t out[10])
{
  HostScene hs;
  build_host_scene(hs, ws, numTlasNodes, numTlasLeaves, numGas, numGasNodes, numGasTris);
  const SceneDesc& sc = hs.sc;

  for (int k = 0; k < 10; ++k) out[k] = 0;
  struct Lane { uint2 smStack[RTC_SM_STACK]; float smRay[RTC_SM_RAY_WORDS]; uint2 lmStack[RTC_LM_STACK]; bool active = false; };
  std::vector<Lane> lanes(32);
  auto simulate = [&](auto proto) {
    using Tr = decltype(proto);
    std::vector<Tr> tr(32);
    for (int l = 0; l < 32; ++l) { tr[l].smStack = lanes[l].smStack; tr[l].smRay = lanes[l].smRay; tr[l].lmStack = lanes[l].lmStack; lanes[l].active = false; }
    uint64_t next = 0;
    for (;;)
    {
      int idle = 0; for (int l = 0; l < 32; ++l) idle += lanes[l].active ? 0 : 1;
      if (idle && next < n && (idle == 32 || (uint32_t)idle >= fetchThreshold))
      {
        out[8]++;
        for (int l = 0; l < 32 && next < n; ++l)
        {
          if (lanes[l].active) continue;
          const rtc_ray& r = rays[next++];
          out[9]++;
          if (tr[l].begin(sc, make_float4(r.ox, r.oy, r.oz, r.tmin), make_float4(r.dx, r.dy, r.dz, r.tmax))) lanes[l].active = true;
        }
      }
      else if (idle == 32) break;
      bool anyNode = false, anyInst = false; uint32_t maxTris = 0; int live = 0;
      for (int l = 0; l < 32; ++l)
      {
        if (!lanes[l].active) continue;
        ++live;
        const uint32_t n0 = tr[l].counts.nodes, t0 = tr[l].counts.tris, i0 = tr[l].counts.insts;
        if (leafThreshold > 0)
        {
          // trace_stream's RTC_LEAF_THRESHOLD path, first half: node visits of the lanes without a pending leaf group
          if (!tr[l].has_leaves()) tr[l].node_phase();
          const uint32_t dnA = tr[l].counts.nodes - n0;
          anyNode |= dnA != 0; out[5] += dnA;
          continue;
        }
        const bool more = tr[l].step(sc);
        const uint32_t dn = tr[l].counts.nodes - n0, dt = tr[l].counts.tris - t0, di = tr[l].counts.insts - i0;
        anyNode |= dn != 0; anyInst |= di != 0; if (dt > maxTris) maxTris = dt;
        out[5] += dn; out[6] += dt; out[7] += di;
        if (!more) lanes[l].active = false;
      }
      if (leafThreshold > 0)
      {
        // second half: the held lanes run their leaf phase together once enough of them hold one
        int nHeld = 0;
        for (int l = 0; l < 32; ++l) if (lanes[l].active && tr[l].has_leaves()) ++nHeld;
        const bool runLeaves = nHeld && ((uint32_t)nHeld >= leafThreshold || nHeld * 4 >= live);
        for (int l = 0; l < 32; ++l)
        {
          if (!lanes[l].active) continue;
          bool running = true;
          const uint32_t t0 = tr[l].counts.tris, i0 = tr[l].counts.insts;
          if (runLeaves && tr[l].has_leaves()) running = tr[l].leaf_phase(sc);
          const uint32_t dt = tr[l].counts.tris - t0, di = tr[l].counts.insts - i0;
          anyInst |= di != 0; if (dt > maxTris) maxTris = dt;
          out[6] += dt; out[7] += di;
          if (running && !tr[l].has_leaves()) running = tr[l].advance(sc);
          if (!running) lanes[l].active = false;
        }
      }
      if (live) { out[0]++; out[1] += anyNode ? 1 : 0; out[2] += maxTris; out[3] += anyInst ? 1 : 0; out[4] += (uint64_t)live; }
    }
  };
  if (any) simulate(Traversal<true, true, 1, false>());
  else     simulate(Traversal<false, true, 1, false>());
  return 0;
}

} // extern "C"
