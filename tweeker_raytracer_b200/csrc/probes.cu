// probes.cu -- roofline denominators the traversal kernels are measured against, taken on the same GPU in the same run
// ("L2 and FP32 peaks are not in MEASURED_PEAKS.json -- measure them with a micro-benchmark in the same run").
//
//   rtc_probe_gather    random 128-bit gathers (LDG.128, one 16-byte record per load, like a node or triangle fetch) from a
//                       working set of `bytes`: L2-resident gather bandwidth for bytes << 126 MB, HBM gather bandwidth above
//   rtc_probe_fp32      dependent-free FFMA streams: FP32 pipe peak (2 flops per FFMA)
//   rtc_probe_issue     FFMA interleaved with LOP3 (fma pipe + alu pipe): warp-instruction issue peak
// All return device time (CUDA events on the context stream) and the work done, so the caller forms the rates.
#include "rtc_internal.h"

namespace {

constexpr int kProbeBlock = 256;

__global__ void __launch_bounds__(kProbeBlock)
k_probe_fill(uint4* __restrict__ buf, uint64_t n)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    buf[i] = make_uint4((uint32_t)i, (uint32_t)(i >> 32), 1u, 2u);
}

// Every thread issues `loads` independent 16-byte loads at pseudo-random record indices (LCG per thread, 8 in flight).
__global__ void __launch_bounds__(kProbeBlock)
k_probe_gather(const uint4* __restrict__ buf, uint32_t mask, uint32_t loads, uint32_t* __restrict__ sink)
{
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  uint32_t acc = 0;
  for (uint32_t i = 0; i < loads; i += 8)
  {
    uint32_t idx[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { s = s * 1664525u + 1013904223u; idx[k] = (s >> 4) & mask; }
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(buf + idx[k]);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += v[k].x ^ v[k].w;
  }
  if (acc == 0x9e3779b9u) *sink = acc;      // never true in practice; keeps the loads alive
}

__global__ void __launch_bounds__(kProbeBlock)
k_probe_fp32(uint32_t iters, float seed, float* __restrict__ sink)
{
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = seed + (float)k;
  const float m = 1.0f + seed * 1e-9f, c = seed * 1e-9f;
  for (uint32_t i = 0; i < iters; ++i)
  {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = fmaf(a[k], m, c);
  }
  float t = 0.0f;
#pragma unroll
  for (int k = 0; k < 8; ++k) t += a[k];
  if (t == 12345.678f) *sink = t;
}

// 64 FFMA + 64 LOP3 per iteration, independent chains, alternating pipes
__global__ void __launch_bounds__(kProbeBlock)
k_probe_issue(uint32_t iters, float seed, float* __restrict__ sink)
{
  float a[8]; uint32_t b[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { a[k] = seed + (float)k; b[k] = __float_as_uint(seed) + 17u * k; }
  const float m = 1.0f + seed * 1e-9f, c = seed * 1e-9f;
  const uint32_t x = __float_as_uint(seed) | 1u, y = ~x;
  for (uint32_t i = 0; i < iters; ++i)
  {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int k = 0; k < 8; ++k)
      {
        a[k] = fmaf(a[k], m, c);
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[k]) : "r"(x), "r"(y));
      }
  }
  float t = 0.0f; uint32_t u = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) { t += a[k]; u ^= b[k]; }
  if (t == 12345.678f && u == 77u) *sink = t;
}

} // namespace

extern "C" int rtc_probe_gather(rtc_context* ctx, uint64_t bytes, uint32_t loadsPerThread, double* gigabytesPerSecond)
{
  if (!ctx || !gigabytesPerSecond) RTC_FAIL("null argument");
  RTC_CUDA(cudaSetDevice(ctx->device));
  uint64_t records = 1;
  while (records * 2 * 16 <= bytes) records *= 2;          // power of two: the index is a mask
  if (records < 1024 || records > (1ull << 32)) RTC_FAIL("working set must be between 16 KB and 64 GB");
  uint4* buf = nullptr; uint32_t* sink = nullptr;
  RTC_CUDA(cudaMalloc(&buf, records * 16));
  if (cudaMalloc(&sink, 4) != cudaSuccess) { cudaFree(buf); RTC_FAIL("cudaMalloc failed"); }
  const int grid = ctx->numSMs * 8;
  k_probe_fill<<<grid, kProbeBlock, 0, ctx->stream>>>(buf, records);
  const uint32_t loads = (loadsPerThread + 7u) & ~7u;
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep)      // the first pass also warms L2
  {
    cudaEventRecord(ctx->evA, ctx->stream);
    k_probe_gather<<<grid, kProbeBlock, 0, ctx->stream>>>(buf, (uint32_t)(records - 1), loads, sink);
    cudaEventRecord(ctx->evB, ctx->stream);
    cudaEventSynchronize(ctx->evB);
    float ms = 0.0f; cudaEventElapsedTime(&ms, ctx->evA, ctx->evB);
    if (rep > 0 && ms < best) best = ms;
  }
  ctx->kernelLaunches += 5;
  cudaFree(buf); cudaFree(sink);
  RTC_CUDA(cudaGetLastError());
  *gigabytesPerSecond = (double)grid * kProbeBlock * loads * 16.0 / (best * 1e-3) / 1e9;
  return 0;
}

// mode 0: FFMA only -> TFLOP/s (2 flops per FFMA); mode 1: FFMA + LOP3 -> warp instructions per second / 1e9
extern "C" int rtc_probe_pipes(rtc_context* ctx, int mode, double* rate)
{
  if (!ctx || !rate) RTC_FAIL("null argument");
  RTC_CUDA(cudaSetDevice(ctx->device));
  float* sink = nullptr;
  RTC_CUDA(cudaMalloc(&sink, 4));
  const int grid = ctx->numSMs * 8;
  const uint32_t iters = 4096;
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep)
  {
    cudaEventRecord(ctx->evA, ctx->stream);
    if (mode == 0) k_probe_fp32<<<grid, kProbeBlock, 0, ctx->stream>>>(iters, 1.0f, sink);
    else           k_probe_issue<<<grid, kProbeBlock, 0, ctx->stream>>>(iters, 1.0f, sink);
    cudaEventRecord(ctx->evB, ctx->stream);
    cudaEventSynchronize(ctx->evB);
    float ms = 0.0f; cudaEventElapsedTime(&ms, ctx->evA, ctx->evB);
    if (rep > 0 && ms < best) best = ms;
  }
  ctx->kernelLaunches += 4;
  cudaFree(sink);
  RTC_CUDA(cudaGetLastError());
  const double threads = (double)grid * kProbeBlock;
  if (mode == 0) *rate = threads * iters * 64.0 * 2.0 / (best * 1e-3) / 1e12;            // TFLOP/s
  else           *rate = threads / 32.0 * iters * 128.0 / (best * 1e-3) / 1e9;            // G warp-instructions/s
  return 0;
}
