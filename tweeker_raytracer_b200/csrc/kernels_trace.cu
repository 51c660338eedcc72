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

// ---- kernels ---------------------------------------------------------------------------------------------------------

template <bool ANY, bool COUNT, class Policy>
__global__ void __launch_bounds__(kTraceBlock, RTC_TRACE_MIN_BLOCKS)
k_trace(const SceneDesc sc, Policy policy, uint32_t n, const uint32_t* __restrict__ nPtr, uint32_t* __restrict__ cursor,
        unsigned long long* __restrict__ counts)
{
  __shared__ uint2 smem[RTC_SM_STACK * kTraceBlock + (11 * kTraceBlock + 1) / 2];     // stack columns, then eleven float columns (trace.cuh smRay)
  const uint32_t count = nPtr ? *nPtr : n;      // the wavefront keeps its queue lengths on the device
  trace_stream<ANY, COUNT, kTraceBlock>(sc, count, cursor, policy, smem, counts);
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

int read_stack_overflows(rtc_context* ctx, uint64_t* out)
{
  unsigned int v = 0;
  RTC_CUDA(cudaMemcpyFromSymbolAsync(&v, g_rtcStackOverflows, sizeof(v), 0, cudaMemcpyDeviceToHost, ctx->stream));
  RTC_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = v;
  return 0;
}
