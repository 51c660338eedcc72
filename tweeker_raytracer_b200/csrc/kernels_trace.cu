// kernels_trace.cu -- the traversal kernels: ray queries (rtc_trace_*), and the wavefront integrator's
// extend (closest hit of the radiance-ray queue) and connect (any hit of the shadow-ray queue).
// Built for sm_100a with FMA contraction ON: only the box tests may contract; the intersector in
// trace.cuh pins its own rounding with intrinsics.
#include "trace.cuh"

namespace {

constexpr int kTraceBlock = 128;

__global__ void __launch_bounds__(kTraceBlock)
k_trace_closest(const SceneDesc sc, const float4* __restrict__ rays, uint64_t n, rtc_hit* __restrict__ hits)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const float4 o = __ldg(rays + 2 * i), d = __ldg(rays + 2 * i + 1);
    TraceHit h;
    trace_ray<false>(sc, o, d, h);
    rtc_hit out; out.t = h.t; out.u = h.u; out.v = h.v; out.inst = h.inst; out.prim = h.prim;
    hits[i] = out;
  }
}

__global__ void __launch_bounds__(kTraceBlock)
k_trace_any(const SceneDesc sc, const float4* __restrict__ rays, uint64_t n, uint32_t* __restrict__ occluded)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const float4 o = __ldg(rays + 2 * i), d = __ldg(rays + 2 * i + 1);
    TraceHit h;
    occluded[i] = trace_ray<true>(sc, o, d, h) ? 1u : 0u;
  }
}

// extend: closest hit for every path id in the queue (raygeneration.cu:84-89 optixTrace RADIANCE)
__global__ void __launch_bounds__(kTraceBlock)
k_extend(const SceneDesc sc, const float4* __restrict__ rayOrg, const float4* __restrict__ rayDir,
         float4* __restrict__ hit, uint32_t* __restrict__ hitInst, const uint32_t* __restrict__ queue, const uint32_t* __restrict__ count)
{
  const uint32_t n = *count;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const uint32_t p = queue[i];
    const float4 o = rayOrg[p], d = rayDir[p];
    TraceHit h;
    trace_ray<false>(sc, o, d, h);
    hit[p] = make_float4(h.t, h.u, h.v, __uint_as_float(h.prim));
    hitInst[p] = h.inst;
  }
}

// connect: visibility of every queued shadow ray (closesthit.cu:281-300 + anyhit.cu:84-91);
// an unoccluded ray adds its pre-multiplied contribution to the path radiance.
__global__ void __launch_bounds__(kTraceBlock)
k_connect(const SceneDesc sc, const float4* __restrict__ shadowOrg, const float4* __restrict__ shadowDir,
          const float4* __restrict__ contrib, float4* __restrict__ radiance, const uint32_t* __restrict__ queue, const uint32_t* __restrict__ count)
{
  const uint32_t n = *count;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const uint32_t p = queue[i];
    const float4 o = shadowOrg[p], d = shadowDir[p];
    TraceHit h;
    if (!trace_ray<true>(sc, o, d, h))
    {
      const float4 c = contrib[p];
      float4 L = radiance[p];
      L.x = __fadd_rn(L.x, c.x); L.y = __fadd_rn(L.y, c.y); L.z = __fadd_rn(L.z, c.z);
      radiance[p] = L;
    }
  }
}

inline int grid_for(const rtc_context* ctx, uint64_t n, int blocksPerSM)
{
  const uint64_t want = (n + kTraceBlock - 1) / kTraceBlock;
  const uint64_t cap = (uint64_t)ctx->numSMs * blocksPerSM;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

} // namespace

int launch_trace_closest(rtc_context* ctx, const SceneDesc* scene, const rtc_ray* rays, uint64_t n, rtc_hit* hits)
{
  if (n == 0) return 0;
  k_trace_closest<<<grid_for(ctx, n, 16), kTraceBlock, 0, ctx->stream>>>(*scene, reinterpret_cast<const float4*>(rays), n, hits);
  ctx->kernelLaunches++;
  RTC_CUDA(cudaGetLastError());
  return 0;
}

int launch_trace_any(rtc_context* ctx, const SceneDesc* scene, const rtc_ray* rays, uint64_t n, uint32_t* occluded)
{
  if (n == 0) return 0;
  k_trace_any<<<grid_for(ctx, n, 16), kTraceBlock, 0, ctx->stream>>>(*scene, reinterpret_cast<const float4*>(rays), n, occluded);
  ctx->kernelLaunches++;
  RTC_CUDA(cudaGetLastError());
  return 0;
}

int launch_extend(rtc_context* ctx, const SceneDesc* scene, const WavefrontBuffers& wf, const uint32_t* queue, const uint32_t* count)
{
  k_extend<<<ctx->numSMs * 16, kTraceBlock, 0, ctx->stream>>>(*scene, wf.rayOrg, wf.rayDir, wf.hit, wf.hitInst, queue, count);
  ctx->kernelLaunches++;
  RTC_CUDA(cudaGetLastError());
  return 0;
}

int launch_connect(rtc_context* ctx, const SceneDesc* scene, const WavefrontBuffers& wf, const uint32_t* count)
{
  k_connect<<<ctx->numSMs * 16, kTraceBlock, 0, ctx->stream>>>(*scene, wf.shadowOrg, wf.shadowDir, wf.shadowContrib, wf.radiance, wf.shadowQueue, count);
  ctx->kernelLaunches++;
  RTC_CUDA(cudaGetLastError());
  return 0;
}
