// kernels_trace.cu -- the traversal kernels: ray queries (rtc_trace_*), and the wavefront integrator's
// extend (closest hit of the radiance-ray queue) and connect (any hit of the shadow-ray queue).
// All of them are persistent-warp kernels over trace_stream() (trace.cuh): one CTA of 128 threads per
// resident slot, rays handed out through a device-side cursor.
// Built for sm_100a with FMA contraction ON: only the box tests may contract; the intersector in
// trace.cuh pins its own rounding with intrinsics.
#include "trace.cuh"

#ifndef RTC_TRACE_MIN_BLOCKS
#define RTC_TRACE_MIN_BLOCKS 8      // resident CTAs per SM the traversal kernels are compiled for (register budget) and launched with
#endif

namespace {

constexpr int kTraceBlock = 128;

// ---- ray sources / hit sinks -------------------------------------------------------------------------------------

// rtc_trace_closest: AoS rays in, rtc_hit out
struct QueryClosest
{
  const float4* __restrict__ rays; rtc_hit* __restrict__ hits;
  __device__ __forceinline__ bool load(uint32_t i, float4& o, float4& d) const { o = __ldg(rays + 2 * (size_t)i); d = __ldg(rays + 2 * (size_t)i + 1); return true; }
  template <class T> __device__ __forceinline__ void store(uint32_t i, const T& tr) const
  {
    const TraceHit h = tr.result();
    rtc_hit out; out.t = h.t; out.u = h.u; out.v = h.v; out.inst = h.inst; out.prim = h.prim;
    hits[i] = out;
  }
};

// rtc_trace_any: AoS rays in, one uint32 per ray out
struct QueryAny
{
  const float4* __restrict__ rays; uint32_t* __restrict__ occluded;
  __device__ __forceinline__ bool load(uint32_t i, float4& o, float4& d) const { o = __ldg(rays + 2 * (size_t)i); d = __ldg(rays + 2 * (size_t)i + 1); return true; }
  template <class T> __device__ __forceinline__ void store(uint32_t i, const T& tr) const { occluded[i] = tr.found() ? 1u : 0u; }
};

// rtc_trace_count: rays in, nothing out (the counters are the result)
struct QueryCount
{
  const float4* __restrict__ rays;
  __device__ __forceinline__ bool load(uint32_t i, float4& o, float4& d) const { o = __ldg(rays + 2 * (size_t)i); d = __ldg(rays + 2 * (size_t)i + 1); return true; }
  template <class T> __device__ __forceinline__ void store(uint32_t, const T&) const {}
};

// extend: queue of path ids -> SoA radiance rays; hit record per path (raygeneration.cu:84-89 optixTrace RADIANCE)
struct ExtendPaths
{
  const uint32_t* __restrict__ queue; const float4* __restrict__ rayOrg; const float4* __restrict__ rayDir;
  float4* __restrict__ hit; uint32_t* __restrict__ hitInst;
  uint32_t path;
  __device__ __forceinline__ bool load(uint32_t i, float4& o, float4& d) { path = queue[i]; o = rayOrg[path]; d = rayDir[path]; return true; }
  template <class T> __device__ __forceinline__ void store(uint32_t, const T& tr) const
  {
    const TraceHit h = tr.result();
    __stcs(hit + path, make_float4(h.t, h.u, h.v, __uint_as_float(h.prim)));
    __stcs(hitInst + path, h.inst);
  }
};

// connect: queue of path ids -> SoA shadow rays (closesthit.cu:281-300 + anyhit.cu:84-91); an unoccluded ray adds its
// pre-multiplied contribution to the path radiance (each path has at most one shadow ray in flight: no race).
struct ConnectPaths
{
  const uint32_t* __restrict__ queue; const float4* __restrict__ shadowOrg; const float4* __restrict__ shadowDir;
  const float4* __restrict__ contrib; float4* __restrict__ radiance;
  uint32_t path;
  __device__ __forceinline__ bool load(uint32_t i, float4& o, float4& d) { path = queue[i]; o = shadowOrg[path]; d = shadowDir[path]; return true; }
  template <class T> __device__ __forceinline__ void store(uint32_t, const T& tr) const
  {
    if (!tr.found())
    {
      const float4 c = __ldcs(contrib + path);
      float4 L = __ldcs(radiance + path);
      L.x = __fadd_rn(L.x, c.x); L.y = __fadd_rn(L.y, c.y); L.z = __fadd_rn(L.z, c.z);
      __stcs(radiance + path, L);
    }
  }
};

// ---- ordered any-hit processing (scenes with cutout materials only; anyhit.cu:46-132) ----------------------------------
// The closest candidate of a ray is found by the normal kernels; when the any-hit program ignores it (k_cutout_radiance /
// k_cutout_shadow in kernels_shade.cu), the ray is traced again for the next candidate AFTER the ignored one.  The ignored
// hit sits in hit[path] / hitInst[path] and is the skip key; tmin moves up to just below its t (ties are resolved by the key).

// radiance ray of a path, again, past the candidate in hit[path]
struct ExtendPathsAfter
{
  const uint32_t* __restrict__ queue; const float4* __restrict__ rayOrg; const float4* __restrict__ rayDir;
  float4* __restrict__ hit; uint32_t* __restrict__ hitInst;
  uint32_t path; float4 prev; uint32_t prevInst;
  __device__ __forceinline__ bool load(uint32_t i, float4& o, float4& d)
  {
    path = queue[i]; o = rayOrg[path]; d = rayDir[path]; prev = hit[path]; prevInst = hitInst[path];
    o.w = fmaxf(o.w, __uint_as_float(__float_as_uint(prev.x) - 1u));      // prev.x > tmin > 0
    return true;
  }
  __device__ __forceinline__ void load_skip(float& t, uint32_t& inst, uint32_t& prim) const { t = prev.x; inst = prevInst; prim = __float_as_uint(prev.w); }
  template <class T> __device__ __forceinline__ void store(uint32_t, const T& tr) const
  {
    const TraceHit h = tr.result();
    hit[path] = make_float4(h.t, h.u, h.v, __uint_as_float(h.prim));
    hitInst[path] = h.inst;
  }
};

// shadow ray of a path as a CLOSEST-hit query (first candidate, or the next one after hit[path] when AFTER)
template <bool AFTER>
struct ConnectClosest
{
  const uint32_t* __restrict__ queue; const float4* __restrict__ shadowOrg; const float4* __restrict__ shadowDir;
  float4* __restrict__ hit; uint32_t* __restrict__ hitInst;
  uint32_t path; float4 prev; uint32_t prevInst;
  __device__ __forceinline__ bool load(uint32_t i, float4& o, float4& d)
  {
    path = queue[i]; o = shadowOrg[path]; d = shadowDir[path];
    if (AFTER) { prev = hit[path]; prevInst = hitInst[path]; o.w = fmaxf(o.w, __uint_as_float(__float_as_uint(prev.x) - 1u)); }
    return true;
  }
  __device__ __forceinline__ void load_skip(float& t, uint32_t& inst, uint32_t& prim) const { t = prev.x; inst = prevInst; prim = __float_as_uint(prev.w); }
  template <class T> __device__ __forceinline__ void store(uint32_t, const T& tr) const
  {
    const TraceHit h = tr.result();
    hit[path] = make_float4(h.t, h.u, h.v, __uint_as_float(h.prim));
    hitInst[path] = h.inst;
  }
};

// ---- kernels ---------------------------------------------------------------------------------------------------------

template <bool ANY, bool COUNT, class Policy, bool SKIP = false>
__global__ void __launch_bounds__(kTraceBlock, SKIP ? 4 : RTC_TRACE_MIN_BLOCKS)
k_trace(const SceneDesc sc, Policy policy, uint32_t n, const uint32_t* __restrict__ nPtr, uint32_t* __restrict__ cursor,
        unsigned long long* __restrict__ counts)
{
  __shared__ uint2 smem[RTC_SM_STACK * kTraceBlock + (11 * kTraceBlock + 1) / 2];     // stack columns, then eleven float columns (trace.cuh smRay)
  const uint32_t count = nPtr ? *nPtr : n;      // the wavefront keeps its queue lengths on the device
  trace_stream<ANY, COUNT, kTraceBlock, SKIP>(sc, count, cursor, policy, smem, counts);
}

inline int persistent_grid(const rtc_context* ctx) { return ctx->numSMs * RTC_TRACE_MIN_BLOCKS; }

} // namespace

int launch_trace_closest(rtc_context* ctx, const SceneDesc* scene, const rtc_ray* rays, uint64_t n, rtc_hit* hits)
{
  if (n == 0) return 0;
  if (n > 0xfffffff0ull) RTC_FAIL("more than 2^32 rays in one call");
  RTC_CUDA(cudaMemsetAsync(ctx->d_cursor, 0, sizeof(uint32_t), ctx->stream));
  QueryClosest p = { reinterpret_cast<const float4*>(rays), hits };
  k_trace<false, false, QueryClosest><<<persistent_grid(ctx), kTraceBlock, 0, ctx->stream>>>(*scene, p, (uint32_t)n, nullptr, ctx->d_cursor, nullptr);
  ctx->kernelLaunches++;
  RTC_CUDA(cudaGetLastError());
  return 0;
}

int launch_trace_any(rtc_context* ctx, const SceneDesc* scene, const rtc_ray* rays, uint64_t n, uint32_t* occluded)
{
  if (n == 0) return 0;
  if (n > 0xfffffff0ull) RTC_FAIL("more than 2^32 rays in one call");
  RTC_CUDA(cudaMemsetAsync(ctx->d_cursor, 0, sizeof(uint32_t), ctx->stream));
  QueryAny p = { reinterpret_cast<const float4*>(rays), occluded };
  k_trace<true, false, QueryAny><<<persistent_grid(ctx), kTraceBlock, 0, ctx->stream>>>(*scene, p, (uint32_t)n, nullptr, ctx->d_cursor, nullptr);
  ctx->kernelLaunches++;
  RTC_CUDA(cudaGetLastError());
  return 0;
}

int launch_trace_count(rtc_context* ctx, const SceneDesc* scene, const rtc_ray* rays, uint64_t n, int anyHit, unsigned long long* d_counts)
{
  if (n == 0) return 0;
  if (n > 0xfffffff0ull) RTC_FAIL("more than 2^32 rays in one call");
  RTC_CUDA(cudaMemsetAsync(ctx->d_cursor, 0, sizeof(uint32_t), ctx->stream));
  QueryCount p = { reinterpret_cast<const float4*>(rays) };
  if (anyHit) k_trace<true, true, QueryCount><<<persistent_grid(ctx), kTraceBlock, 0, ctx->stream>>>(*scene, p, (uint32_t)n, nullptr, ctx->d_cursor, d_counts);
  else        k_trace<false, true, QueryCount><<<persistent_grid(ctx), kTraceBlock, 0, ctx->stream>>>(*scene, p, (uint32_t)n, nullptr, ctx->d_cursor, d_counts);
  ctx->kernelLaunches++;
  RTC_CUDA(cudaGetLastError());
  return 0;
}

int launch_extend(rtc_context* ctx, const SceneDesc* scene, const WavefrontBuffers& wf, const uint32_t* queue, const uint32_t* count,
                  uint32_t* cursor, bool countWork)
{
  if (int rc = profile_begin(ctx, RTC_KERNEL_EXTEND)) return rc;
  ExtendPaths p = { queue, wf.rayOrg, wf.rayDir, wf.hit, wf.hitInst, 0u };
  if (countWork) k_trace<false, true, ExtendPaths><<<persistent_grid(ctx), kTraceBlock, 0, ctx->stream>>>(*scene, p, 0u, count, cursor, ctx->d_launchCounts);
  else           k_trace<false, false, ExtendPaths><<<persistent_grid(ctx), kTraceBlock, 0, ctx->stream>>>(*scene, p, 0u, count, cursor, nullptr);
  ctx->kernelLaunches++;
  RTC_CUDA(cudaGetLastError());
  return profile_end(ctx);
}

int launch_connect(rtc_context* ctx, const SceneDesc* scene, const WavefrontBuffers& wf, const uint32_t* count, uint32_t* cursor, bool countWork)
{
  if (int rc = profile_begin(ctx, RTC_KERNEL_CONNECT)) return rc;
  ConnectPaths p = { wf.shadowQueue, wf.shadowOrg, wf.shadowDir, wf.shadowContrib, wf.radiance, 0u };
  if (countWork) k_trace<true, true, ConnectPaths><<<persistent_grid(ctx), kTraceBlock, 0, ctx->stream>>>(*scene, p, 0u, count, cursor, ctx->d_launchCounts + 4);
  else           k_trace<true, false, ConnectPaths><<<persistent_grid(ctx), kTraceBlock, 0, ctx->stream>>>(*scene, p, 0u, count, cursor, nullptr);
  ctx->kernelLaunches++;
  RTC_CUDA(cudaGetLastError());
  return profile_end(ctx);
}

// Re-trace of the radiance rays in `queue` past their ignored candidate (see ExtendPathsAfter).
int launch_extend_after(rtc_context* ctx, const SceneDesc* scene, const WavefrontBuffers& wf, const uint32_t* queue, const uint32_t* count, uint32_t* cursor)
{
  if (int rc = profile_begin(ctx, RTC_KERNEL_EXTEND)) return rc;
  ExtendPathsAfter p = { queue, wf.rayOrg, wf.rayDir, wf.hit, wf.hitInst, 0u, make_float4(0.f, 0.f, 0.f, 0.f), 0u };
  RTC_CUDA(cudaMemsetAsync(cursor, 0, sizeof(uint32_t), ctx->stream));
  k_trace<false, false, ExtendPathsAfter, true><<<persistent_grid(ctx), kTraceBlock, 0, ctx->stream>>>(*scene, p, 0u, count, cursor, nullptr);
  ctx->kernelLaunches++;
  RTC_CUDA(cudaGetLastError());
  return profile_end(ctx);
}

// Shadow rays of `queue` as closest-hit queries into hit[] / hitInst[] (after = past the ignored candidate already there).
int launch_connect_closest(rtc_context* ctx, const SceneDesc* scene, const WavefrontBuffers& wf, const uint32_t* queue, const uint32_t* count,
                           uint32_t* cursor, bool after)
{
  if (int rc = profile_begin(ctx, RTC_KERNEL_CONNECT)) return rc;
  RTC_CUDA(cudaMemsetAsync(cursor, 0, sizeof(uint32_t), ctx->stream));
  if (after)
  {
    ConnectClosest<true> p = { queue, wf.shadowOrg, wf.shadowDir, wf.hit, wf.hitInst, 0u, make_float4(0.f, 0.f, 0.f, 0.f), 0u };
    k_trace<false, false, ConnectClosest<true>, true><<<persistent_grid(ctx), kTraceBlock, 0, ctx->stream>>>(*scene, p, 0u, count, cursor, nullptr);
  }
  else
  {
    ConnectClosest<false> p = { queue, wf.shadowOrg, wf.shadowDir, wf.hit, wf.hitInst, 0u, make_float4(0.f, 0.f, 0.f, 0.f), 0u };
    k_trace<false, false, ConnectClosest<false>, false><<<persistent_grid(ctx), kTraceBlock, 0, ctx->stream>>>(*scene, p, 0u, count, cursor, nullptr);
  }
  ctx->kernelLaunches++;
  RTC_CUDA(cudaGetLastError());
  return profile_end(ctx);
}

int read_stack_overflows(rtc_context* ctx, uint64_t* out)
{
  unsigned int v = 0;
  RTC_CUDA(cudaMemcpyFromSymbolAsync(&v, g_rtcStackOverflows, sizeof(v), 0, cudaMemcpyDeviceToHost, ctx->stream));
  RTC_CUDA(cudaStreamSynchronize(ctx->stream));
  uint64_t primary = 0;
  if (int rc = read_stack_overflows_primary(ctx, &primary)) return rc;
  *out = (uint64_t)v + primary;
  return 0;
}
