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

// counting variant of the two query kernels: same traversal, plus per-ray work counters reduced per warp
template <bool ANY>
__global__ void __launch_bounds__(kTraceBlock)
k_trace_count(const SceneDesc sc, const float4* __restrict__ rays, uint64_t n, unsigned long long* __restrict__ counts)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long nodes = 0, tris = 0, insts = 0, nrays = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const float4 o = __ldg(rays + 2 * i), d = __ldg(rays + 2 * i + 1);
    TraceHit h; TraceCounts c = { 0u, 0u, 0u };
    trace_ray<ANY, true>(sc, o, d, h, &c);
    nodes += c.nodes; tris += c.tris; insts += c.insts; nrays += 1;
  }
  for (int off = 16; off; off >>= 1)
  {
    nodes += __shfl_down_sync(0xffffffffu, nodes, off); tris += __shfl_down_sync(0xffffffffu, tris, off);
    insts += __shfl_down_sync(0xffffffffu, insts, off); nrays += __shfl_down_sync(0xffffffffu, nrays, off);
  }
  if ((threadIdx.x & 31) == 0)
  {
    atomicAdd(counts + 0, nodes); atomicAdd(counts + 1, tris); atomicAdd(counts + 2, insts); atomicAdd(counts + 3, nrays);
  }
}

// warp-reduces per-thread work counters and adds them to counts[0..3]
__device__ __forceinline__ void flush_counts(unsigned long long nodes, unsigned long long tris, unsigned long long insts, unsigned long long nrays,
                                             unsigned long long* __restrict__ counts)
{
  for (int off = 16; off; off >>= 1)
  {
    nodes += __shfl_down_sync(0xffffffffu, nodes, off); tris += __shfl_down_sync(0xffffffffu, tris, off);
    insts += __shfl_down_sync(0xffffffffu, insts, off); nrays += __shfl_down_sync(0xffffffffu, nrays, off);
  }
  if ((threadIdx.x & 31) == 0 && nrays)
  {
    atomicAdd(counts + 0, nodes); atomicAdd(counts + 1, tris); atomicAdd(counts + 2, insts); atomicAdd(counts + 3, nrays);
  }
}

// extend: closest hit for every path id in the queue (raygeneration.cu:84-89 optixTrace RADIANCE)
template <bool COUNT>
__global__ void __launch_bounds__(kTraceBlock)
k_extend(const SceneDesc sc, const float4* __restrict__ rayOrg, const float4* __restrict__ rayDir,
         float4* __restrict__ hit, uint32_t* __restrict__ hitInst, const uint32_t* __restrict__ queue, const uint32_t* __restrict__ count,
         unsigned long long* __restrict__ counts)
{
  const uint32_t n = *count;
  const uint32_t stride = gridDim.x * blockDim.x;
  unsigned long long nodes = 0, tris = 0, insts = 0, nrays = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const uint32_t p = queue[i];
    const float4 o = rayOrg[p], d = rayDir[p];
    TraceHit h; TraceCounts c = { 0u, 0u, 0u };
    trace_ray<false, COUNT>(sc, o, d, h, &c);
    hit[p] = make_float4(h.t, h.u, h.v, __uint_as_float(h.prim));
    hitInst[p] = h.inst;
    if (COUNT) { nodes += c.nodes; tris += c.tris; insts += c.insts; nrays += 1; }
  }
  if (COUNT) flush_counts(nodes, tris, insts, nrays, counts);
}

// connect: visibility of every queued shadow ray (closesthit.cu:281-300 + anyhit.cu:84-91);
// an unoccluded ray adds its pre-multiplied contribution to the path radiance.
template <bool COUNT>
__global__ void __launch_bounds__(kTraceBlock)
k_connect(const SceneDesc sc, const float4* __restrict__ shadowOrg, const float4* __restrict__ shadowDir,
          const float4* __restrict__ contrib, float4* __restrict__ radiance, const uint32_t* __restrict__ queue, const uint32_t* __restrict__ count,
          unsigned long long* __restrict__ counts)
{
  const uint32_t n = *count;
  const uint32_t stride = gridDim.x * blockDim.x;
  unsigned long long nodes = 0, tris = 0, insts = 0, nrays = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const uint32_t p = queue[i];
    const float4 o = shadowOrg[p], d = shadowDir[p];
    TraceHit h; TraceCounts c = { 0u, 0u, 0u };
    if (!trace_ray<true, COUNT>(sc, o, d, h, &c))
    {
      const float4 cc = contrib[p];
      float4 L = radiance[p];
      L.x = __fadd_rn(L.x, cc.x); L.y = __fadd_rn(L.y, cc.y); L.z = __fadd_rn(L.z, cc.z);
      radiance[p] = L;
    }
    if (COUNT) { nodes += c.nodes; tris += c.tris; insts += c.insts; nrays += 1; }
  }
  if (COUNT) flush_counts(nodes, tris, insts, nrays, counts);
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

int launch_trace_count(rtc_context* ctx, const SceneDesc* scene, const rtc_ray* rays, uint64_t n, int anyHit, unsigned long long* d_counts)
{
  if (n == 0) return 0;
  if (anyHit) k_trace_count<true><<<grid_for(ctx, n, 16), kTraceBlock, 0, ctx->stream>>>(*scene, reinterpret_cast<const float4*>(rays), n, d_counts);
  else        k_trace_count<false><<<grid_for(ctx, n, 16), kTraceBlock, 0, ctx->stream>>>(*scene, reinterpret_cast<const float4*>(rays), n, d_counts);
  ctx->kernelLaunches++;
  RTC_CUDA(cudaGetLastError());
  return 0;
}

int launch_extend(rtc_context* ctx, const SceneDesc* scene, const WavefrontBuffers& wf, const uint32_t* queue, const uint32_t* count, bool countWork)
{
  if (int rc = profile_begin(ctx, RTC_KERNEL_EXTEND)) return rc;
  if (countWork) k_extend<true><<<ctx->numSMs * 16, kTraceBlock, 0, ctx->stream>>>(*scene, wf.rayOrg, wf.rayDir, wf.hit, wf.hitInst, queue, count, ctx->d_launchCounts);
  else           k_extend<false><<<ctx->numSMs * 16, kTraceBlock, 0, ctx->stream>>>(*scene, wf.rayOrg, wf.rayDir, wf.hit, wf.hitInst, queue, count, nullptr);
  ctx->kernelLaunches++;
  RTC_CUDA(cudaGetLastError());
  return profile_end(ctx);
}

int launch_connect(rtc_context* ctx, const SceneDesc* scene, const WavefrontBuffers& wf, const uint32_t* count, bool countWork)
{
  if (int rc = profile_begin(ctx, RTC_KERNEL_CONNECT)) return rc;
  if (countWork) k_connect<true><<<ctx->numSMs * 16, kTraceBlock, 0, ctx->stream>>>(*scene, wf.shadowOrg, wf.shadowDir, wf.shadowContrib, wf.radiance, wf.shadowQueue, count, ctx->d_launchCounts + 4);
  else           k_connect<false><<<ctx->numSMs * 16, kTraceBlock, 0, ctx->stream>>>(*scene, wf.shadowOrg, wf.shadowDir, wf.shadowContrib, wf.radiance, wf.shadowQueue, count, nullptr);
  ctx->kernelLaunches++;
  RTC_CUDA(cudaGetLastError());
  return profile_end(ctx);
}
