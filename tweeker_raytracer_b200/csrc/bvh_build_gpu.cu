// bvh_build_gpu.cu -- device builder for geometry acceleration structures (replaces optixAccelBuild over
// OPTIX_BUILD_INPUT_TYPE_TRIANGLES, apps/rtigo3/src/Device.cpp:1391-1405, for inputs too large for the host builder).
//
//   1. triangle bounds + scene bounds                                   k_tri_bounds
//   2. 63-bit Morton codes of the box centres, radix sort               k_morton, cub::DeviceRadixSort (library sort; build path only)
//   3. binary radix tree over the sorted codes (Karras 2012)            k_radix_tree
//   4. bottom-up box fit                                                k_fit_boxes
//   5. SAH-binned top levels: the subtrees of a cut of ~4096 nodes are re-linked by a binned-SAH
//      build over their boxes (the cut is tiny, so it runs on the host) k_mark_cut, build_binary_sah_host
//   6. collapse to 8-wide nodes with quantised child boxes, level by    k_collapse_level
//      level, emitting triangles in leaf order
// Everything stays in device memory; the result is written straight into GasRecord::d_nodes / d_tris.
// The node format and the conservative quantisation rules are those of bvh_build_host.cpp.
#include "rtc_internal.h"

#include <cub/device/device_radix_sort.cuh>

#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace {

constexpr int kB = 256;
constexpr uint32_t kLeafMax = 3;      // a wide leaf child holds 1..3 triangles

// ---- float atomics through the order-preserving integer image -------------------------------------------------
__device__ __forceinline__ void atomicMinF(float* addr, float v)
{
  if (v >= 0.0f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else           atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomicMaxF(float* addr, float v)
{
  if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else           atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

struct BuildArrays
{
  // per triangle (input order)
  float4* triLo; float4* triHi;
  // sort
  unsigned long long* keys; unsigned long long* keysAlt; uint32_t* order; uint32_t* orderAlt;
  // binary tree: node ids 0..n-2 internal, n-1..2n-2 leaves (sorted position p -> n-1+p); ids >= 2n-1: SAH top nodes
  int2* child;            // children of internal / top nodes
  int*  parent;
  int2* range;            // sorted positions [first, last] covered by a node
  float4* boxLo; float4* boxHi;
  uint32_t* flags;        // arrival counters of the bottom-up fit
  float* sceneBox;        // lo[3], hi[3]
};

__global__ void __launch_bounds__(kB)
k_tri_bounds(const uint8_t* __restrict__ verts, uint32_t stride, const uint32_t* __restrict__ idx, uint32_t n, uint32_t numVerts,
             float4* __restrict__ triLo, float4* __restrict__ triHi, float* __restrict__ sceneBox, uint32_t* __restrict__ badIndex)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
  if (t < n)
  {
    for (int c = 0; c < 3; ++c)
    {
      const uint32_t vi = __ldg(idx + 3u * (size_t)t + c);
      if (vi >= numVerts) { atomicExch(badIndex, 1u); continue; }
      const float* p = reinterpret_cast<const float*>(verts + (size_t)vi * stride);
      for (int k = 0; k < 3; ++k) { const float v = __ldg(p + k); lo[k] = fminf(lo[k], v); hi[k] = fmaxf(hi[k], v); }
    }
    triLo[t] = make_float4(lo[0], lo[1], lo[2], 0.0f);
    triHi[t] = make_float4(hi[0], hi[1], hi[2], 0.0f);
  }
  // block reduction of the scene box, then one atomic per block and plane
  __shared__ float sLo[3][kB / 32], sHi[3][kB / 32];
  for (int k = 0; k < 3; ++k)
  {
    float a = lo[k], b = hi[k];
    for (int off = 16; off; off >>= 1) { a = fminf(a, __shfl_down_sync(0xffffffffu, a, off)); b = fmaxf(b, __shfl_down_sync(0xffffffffu, b, off)); }
    if ((threadIdx.x & 31) == 0) { sLo[k][threadIdx.x >> 5] = a; sHi[k][threadIdx.x >> 5] = b; }
  }
  __syncthreads();
  if (threadIdx.x < 3)
  {
    float a = FLT_MAX, b = -FLT_MAX;
    for (int w = 0; w < kB / 32; ++w) { a = fminf(a, sLo[threadIdx.x][w]); b = fmaxf(b, sHi[threadIdx.x][w]); }
    if (a <= b) { atomicMinF(sceneBox + threadIdx.x, a); atomicMaxF(sceneBox + 3 + threadIdx.x, b); }
  }
}

__device__ __forceinline__ unsigned long long spread21(uint32_t v)
{
  unsigned long long x = v & 0x1fffffull;
  x = (x | (x << 32)) & 0x1f00000000ffffull;
  x = (x | (x << 16)) & 0x1f0000ff0000ffull;
  x = (x | (x << 8))  & 0x100f00f00f00f00full;
  x = (x | (x << 4))  & 0x10c30c30c30c30c3ull;
  x = (x | (x << 2))  & 0x1249249249249249ull;
  return x;
}

__global__ void __launch_bounds__(kB)
k_morton(const float4* __restrict__ triLo, const float4* __restrict__ triHi, uint32_t n, const float* __restrict__ sceneBox,
         unsigned long long* __restrict__ keys, uint32_t* __restrict__ order)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const float4 lo = triLo[t], hi = triHi[t];
  const float c[3] = { 0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z) };
  uint32_t q[3];
  for (int k = 0; k < 3; ++k)
  {
    const float ext = sceneBox[3 + k] - sceneBox[k];
    float u = ext > 0.0f ? (c[k] - sceneBox[k]) / ext : 0.5f;
    u = fminf(fmaxf(u, 0.0f), 1.0f);
    q[k] = (uint32_t)fminf(u * 2097152.0f, 2097151.0f);
  }
  keys[t] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
  order[t] = t;
}

// common prefix length of sorted keys i and j, index bits break ties between equal codes
__device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, int n, int i, int j)
{
  if (j < 0 || j >= n) return -1;
  const unsigned long long a = keys[i], b = keys[j];
  if (a != b) return __clzll((long long)(a ^ b));
  return 64 + __clz(i ^ j);
}

__global__ void __launch_bounds__(kB)
k_radix_tree(const unsigned long long* __restrict__ keys, int n, int2* __restrict__ child, int* __restrict__ parent, int2* __restrict__ range)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = delta(keys, n, i, i - d);
  int lmax = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1) if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = delta(keys, n, i, j);
  int s = 0;
  for (int t = (l + 1) >> 1; ; t = (t + 1) >> 1)
  {
    if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    if (t == 1) break;
  }
  const int gamma = i + s * d + min(d, 0);
  const int first = min(i, j), last = max(i, j);
  const int left = (first == gamma) ? (n - 1 + gamma) : gamma;
  const int right = (last == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
  child[i] = make_int2(left, right);
  range[i] = make_int2(first, last);
  parent[left] = i;
  parent[right] = i;
  if (i == 0) parent[0] = -1;
}

__global__ void __launch_bounds__(kB)
k_fit_boxes(int n, const uint32_t* __restrict__ order, const float4* __restrict__ triLo, const float4* __restrict__ triHi,
            const int2* __restrict__ child, const int* __restrict__ parent, float4* __restrict__ boxLo, float4* __restrict__ boxHi,
            uint32_t* __restrict__ flags)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint32_t t = order[p];
  float4 lo = triLo[t], hi = triHi[t];
  int node = n - 1 + p;
  boxLo[node] = lo; boxHi[node] = hi;
  if (n == 1) return;
  __threadfence();
  node = parent[node];
  while (node >= 0)
  {
    if (atomicAdd(flags + node, 1u) == 0u) return;      // the second child to arrive carries on
    __threadfence();
    const int2 c = child[node];
    const volatile float4* vLo = boxLo; const volatile float4* vHi = boxHi;
    const float lx = fminf(vLo[c.x].x, vLo[c.y].x), ly = fminf(vLo[c.x].y, vLo[c.y].y), lz = fminf(vLo[c.x].z, vLo[c.y].z);
    const float hx = fmaxf(vHi[c.x].x, vHi[c.y].x), hy = fmaxf(vHi[c.x].y, vHi[c.y].y), hz = fmaxf(vHi[c.x].z, vHi[c.y].z);
    boxLo[node] = make_float4(lx, ly, lz, 0.0f); boxHi[node] = make_float4(hx, hy, hz, 0.0f);
    __threadfence();
    node = parent[node];
  }
}

// nodes of the cut: subtree size <= limit while the parent's is larger
__global__ void __launch_bounds__(kB)
k_mark_cut(int n, int limit, const int* __restrict__ parent, const int2* __restrict__ range, int* __restrict__ cutNodes, uint32_t* __restrict__ cutCount, uint32_t cutCapacity)
{
  const int node = blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= 2 * n - 1 || node == 0) return;
  const int size = (node < n - 1) ? (range[node].y - range[node].x + 1) : 1;
  const int2 pr = range[parent[node]];
  if (size <= limit && (pr.y - pr.x + 1) > limit)
  {
    const uint32_t slot = atomicAdd(cutCount, 1u);
    if (slot < cutCapacity) cutNodes[slot] = node;
  }
}

__global__ void k_gather_boxes(const int* __restrict__ nodes, uint32_t count, const float4* __restrict__ boxLo, const float4* __restrict__ boxHi, PrimBox* __restrict__ out)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float4 lo = boxLo[nodes[i]], hi = boxHi[nodes[i]];
  PrimBox b; b.lo[0] = lo.x; b.lo[1] = lo.y; b.lo[2] = lo.z; b.hi[0] = hi.x; b.hi[1] = hi.y; b.hi[2] = hi.z;
  out[i] = b;
}

struct CollapseParams
{
  int n;                      // triangles
  const int2* child; const int2* range; const float4* boxLo; const float4* boxHi;
  const uint32_t* order;      // sorted position -> triangle
  const uint8_t* verts; uint32_t stride; const uint32_t* idx;
  Node8* nodes; float4* tris;
  uint32_t* nodeCount; uint32_t* triCount;
  int leafMax;                // triangles per leaf child, 1..kLeafMax
};

__device__ __forceinline__ float half_area(const float4 lo, const float4 hi)
{
  const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
  return dx * dy + dy * dz + dz * dx;
}
__device__ __forceinline__ float pow2f_dev(int e) { return __uint_as_float((uint32_t)(e + 127) << 23); }

// One thread = one wide node of the current level: opens the largest children of its binary subtree until eight remain,
// assigns octant slots, quantises conservatively, allocates its children contiguously and queues them for the next level.
__global__ void __launch_bounds__(128)
k_collapse_level(const CollapseParams P, const int2* __restrict__ work, uint32_t workCount, int2* __restrict__ nextWork, uint32_t* __restrict__ nextCount)
{
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= workCount) return;
  const int src = work[w].x;
  const uint32_t dst = (uint32_t)work[w].y;
  const int n = P.n;
  auto size_of = [&](int node) { return (node >= n - 1 && node < 2 * n - 1) ? 1 : (P.range[node].y - P.range[node].x + 1); };
  auto is_leaf = [&](int node) { return size_of(node) <= P.leafMax; };

  int kids[8]; int nk = 0;
  if (is_leaf(src)) kids[nk++] = src;
  else { const int2 c = P.child[src]; kids[nk++] = c.x; kids[nk++] = c.y; }
  while (nk < 8)
  {
    int best = -1; float bestArea = -1.0f;
    for (int i = 0; i < nk; ++i)
      if (!is_leaf(kids[i])) { const float a = half_area(P.boxLo[kids[i]], P.boxHi[kids[i]]); if (a > bestArea) { bestArea = a; best = i; } }
    if (best < 0) break;
    const int2 c = P.child[kids[best]];
    kids[best] = c.x; kids[nk++] = c.y;
  }

  const float4 blo = P.boxLo[src], bhi = P.boxHi[src];
  const float boxLo[3] = { blo.x, blo.y, blo.z }, boxHi[3] = { bhi.x, bhi.y, bhi.z };
  float kLo[8][3], kHi[8][3];
  for (int i = 0; i < nk; ++i)
  {
    const float4 a = P.boxLo[kids[i]], b = P.boxHi[kids[i]];
    kLo[i][0] = a.x; kLo[i][1] = a.y; kLo[i][2] = a.z; kHi[i][0] = b.x; kHi[i][1] = b.y; kHi[i][2] = b.z;
  }

  // octant slots: greedy maximum of dot(child centre - node centre, slot sign vector)
  int kidAt[8]; for (int s = 0; s < 8; ++s) kidAt[s] = -1;
  {
    float centre[3]; for (int k = 0; k < 3; ++k) centre[k] = 0.5f * (boxLo[k] + boxHi[k]);
    uint32_t slotUsed = 0, kidDone = 0;
    for (int round = 0; round < nk; ++round)
    {
      int bi = -1, bs = -1; float bc = -FLT_MAX;
      for (int i = 0; i < nk; ++i)
      {
        if (kidDone & (1u << i)) continue;
        const float d0 = 0.5f * (kLo[i][0] + kHi[i][0]) - centre[0], d1 = 0.5f * (kLo[i][1] + kHi[i][1]) - centre[1], d2 = 0.5f * (kLo[i][2] + kHi[i][2]) - centre[2];
        for (int s = 0; s < 8; ++s)
        {
          if (slotUsed & (1u << s)) continue;
          const float c = ((s & 4) ? d0 : -d0) + ((s & 2) ? d1 : -d1) + ((s & 1) ? d2 : -d2);
          if (c > bc || bi < 0) { bc = c; bi = i; bs = s; }
        }
      }
      kidAt[bs] = bi; slotUsed |= 1u << bs; kidDone |= 1u << bi;
    }
  }

  // quantisation frame (same rules as bvh_build_host.cpp): grid step 2^e with 254 steps covering the box, origin a
  // sixteenth of a step below it
  float maxExt = 0.0f; for (int k = 0; k < 3; ++k) maxExt = fmaxf(maxExt, boxHi[k] - boxLo[k]);
  int e[3]; float p[3], step[3];
  for (int k = 0; k < 3; ++k)
  {
    const float ext = fmaxf(boxHi[k] - boxLo[k], fmaxf(maxExt * 0x1p-20f, 1.0e-30f));
    int ek; frexpf(ext / 253.0f, &ek);
    ek = min(max(ek, -120), 120);
    for (;;)
    {
      step[k] = pow2f_dev(ek);
      p[k] = boxLo[k] - step[k] * 0.0625f;
      if (!(p[k] < boxLo[k])) p[k] = nextafterf(boxLo[k], -INFINITY);
      if (fmaf(254.0f, step[k], p[k]) >= boxHi[k] || ek >= 120) break;
      ++ek;
    }
    e[k] = ek;
  }

  Node8 node;
  node.px = p[0]; node.py = p[1]; node.pz = p[2];
  node.ex = (uint8_t)(e[0] + 127); node.ey = (uint8_t)(e[1] + 127); node.ez = (uint8_t)(e[2] + 127);
  node.imask = 0;
  uint32_t numInner = 0, numLeafTris = 0;
  for (int s = 0; s < 8; ++s)
  {
    node.meta[s] = 0;
    if (kidAt[s] < 0) continue;
    const int kid = kids[kidAt[s]];
    if (!is_leaf(kid)) { node.imask |= (uint8_t)(1u << s); ++numInner; }
    else numLeafTris += (uint32_t)size_of(kid);
  }
  node.childBase = numInner ? atomicAdd(P.nodeCount, numInner) : 0u;
  node.triBase = numLeafTris ? atomicAdd(P.triCount, numLeafTris) : 0u;
  const uint32_t nextBase = numInner ? atomicAdd(nextCount, numInner) : 0u;

  uint8_t* qlo[3] = { node.qlox, node.qloy, node.qloz };
  uint8_t* qhi[3] = { node.qhix, node.qhiy, node.qhiz };
  uint32_t triOffset = 0, rel = 0;
  for (int s = 0; s < 8; ++s)
  {
    if (kidAt[s] < 0) { for (int k = 0; k < 3; ++k) { qlo[k][s] = 255; qhi[k][s] = 0; } continue; }
    const int ki = kidAt[s], kid = kids[ki];
    for (int k = 0; k < 3; ++k)
    {
      int ql = (int)floor(((double)kLo[ki][k] - (double)p[k]) / (double)step[k]);
      int qh = (int)ceil(((double)kHi[ki][k] - (double)p[k]) / (double)step[k]);
      ql = min(max(ql, 0), 255); qh = min(max(qh, 0), 255);
      while (ql > 0 && !(fmaf((float)ql, step[k], p[k]) <= kLo[ki][k] - step[k] * 0.015625f)) --ql;
      while (qh < 255 && !(fmaf((float)qh, step[k], p[k]) >= kHi[ki][k] + step[k] * 0.015625f)) ++qh;
      qlo[k][s] = (uint8_t)ql; qhi[k][s] = (uint8_t)qh;
    }
    if (node.imask & (1u << s))
    {
      nextWork[nextBase + rel] = make_int2(kid, (int)(node.childBase + rel));
      ++rel;
    }
    else
    {
      const int cnt = size_of(kid);
      const int first = (kid >= n - 1 && kid < 2 * n - 1) ? kid - (n - 1) : P.range[kid].x;
      node.meta[s] = (uint8_t)(((uint32_t)cnt << 5) | triOffset);
      for (int i = 0; i < cnt; ++i)
      {
        const uint32_t prim = P.order[first + i];
        float4* out = P.tris + (size_t)(node.triBase + triOffset + (uint32_t)i) * 3u;
        for (int c = 0; c < 3; ++c)
        {
          const uint32_t vi = __ldg(P.idx + 3u * (size_t)prim + c);
          const float* v = reinterpret_cast<const float*>(P.verts + (size_t)vi * P.stride);
          out[c] = make_float4(__ldg(v), __ldg(v + 1), __ldg(v + 2), c == 0 ? __uint_as_float(prim) : 0.0f);
        }
      }
      triOffset += (uint32_t)cnt;
    }
  }
  P.nodes[dst] = node;
}

// Build scratch comes from the stream-ordered pool (cudaMallocAsync): a synchronous cudaMalloc/cudaFree of a dozen
// multi-gigabyte arrays costs far more wall time than the build kernels themselves (100 M triangles: 57 ms of kernels).
struct DeviceFree
{
  cudaStream_t stream = nullptr;
  std::vector<void*> ptrs;
  ~DeviceFree() { for (void* p : ptrs) cudaFreeAsync(p, stream); }
  template <class T> cudaError_t alloc(T** p, size_t count)
  {
    cudaError_t e = cudaMallocAsync((void**)p, count * sizeof(T) + 16, stream);
    if (e == cudaSuccess) ptrs.push_back(*p);
    return e;
  }
};

} // namespace

// Binned-SAH binary build over a few thousand boxes on the host (bvh_build_host.cpp); returns nodes as (left, right)
// with negative values ~leafIndex, root first.
void build_binary_sah_host(const PrimBox* prims, uint32_t numPrims, std::vector<int2>& children, std::vector<PrimBox>& boxes);

static int build_gas_gpu_impl(rtc_context* ctx, GasRecord& rec)
{
  const uint32_t n = rec.numTris;
  if (n == 0) RTC_FAIL("build_gas_gpu needs at least one triangle");
  if (n > 0x3fffffffu) RTC_FAIL("too many triangles for one GAS");
  cudaStream_t st = ctx->stream;
  const bool verbose = std::getenv("RTC_BUILD_VERBOSE") != nullptr;     // phase timings on stderr
  auto tick = std::chrono::steady_clock::now();
  auto phase = [&](const char* name) {
    if (!verbose) return;
    cudaStreamSynchronize(st);
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "build_gas_gpu[%u tris] %-10s %8.3f ms\n", n, name, std::chrono::duration<double, std::milli>(now - tick).count());
    tick = now;
  };
  DeviceFree scratch;
  scratch.stream = st;
  BuildArrays A{};
  uint32_t* d_counters = nullptr;       // [0] bad index, [1] cut count, [2] wide nodes, [3] tris, [4..5] level queue counts
  int* d_cut = nullptr; PrimBox* d_cutBoxes = nullptr;
  const uint32_t cutCapacity = 1u << 16;
  const size_t numBinary = 2 * (size_t)n - 1;
  const size_t topCapacity = 2 * (size_t)cutCapacity;      // SAH top nodes appended behind the radix tree
  RTC_CUDA(scratch.alloc(&A.triLo, n)); RTC_CUDA(scratch.alloc(&A.triHi, n));
  RTC_CUDA(scratch.alloc(&A.keys, n)); RTC_CUDA(scratch.alloc(&A.keysAlt, n));
  RTC_CUDA(scratch.alloc(&A.order, n)); RTC_CUDA(scratch.alloc(&A.orderAlt, n));
  RTC_CUDA(scratch.alloc(&A.child, numBinary + topCapacity)); RTC_CUDA(scratch.alloc(&A.parent, numBinary));
  RTC_CUDA(scratch.alloc(&A.range, numBinary + topCapacity));
  RTC_CUDA(scratch.alloc(&A.boxLo, numBinary + topCapacity)); RTC_CUDA(scratch.alloc(&A.boxHi, numBinary + topCapacity));
  RTC_CUDA(scratch.alloc(&A.flags, n)); RTC_CUDA(scratch.alloc(&A.sceneBox, 8));
  RTC_CUDA(scratch.alloc(&d_counters, 8)); RTC_CUDA(scratch.alloc(&d_cut, cutCapacity)); RTC_CUDA(scratch.alloc(&d_cutBoxes, cutCapacity));

  phase("alloc");
  const float initBox[6] = { FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX };
  RTC_CUDA(cudaMemcpyAsync(A.sceneBox, initBox, sizeof(initBox), cudaMemcpyHostToDevice, st));
  RTC_CUDA(cudaMemsetAsync(d_counters, 0, 8 * sizeof(uint32_t), st));
  RTC_CUDA(cudaMemsetAsync(A.flags, 0, (size_t)n * sizeof(uint32_t), st));
  const unsigned gridN = (n + kB - 1) / kB;
  k_tri_bounds<<<gridN, kB, 0, st>>>((const uint8_t*)(uintptr_t)rec.attributes, rec.strideBytes, (const uint32_t*)(uintptr_t)rec.indices, n, rec.numVerts,
                                      A.triLo, A.triHi, A.sceneBox, d_counters);
  k_morton<<<gridN, kB, 0, st>>>(A.triLo, A.triHi, n, A.sceneBox, A.keys, A.order);
  ctx->kernelLaunches += 2;
  {
    cub::DoubleBuffer<unsigned long long> k(A.keys, A.keysAlt);
    cub::DoubleBuffer<uint32_t> v(A.order, A.orderAlt);
    size_t tempBytes = 0;
    RTC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, k, v, (int)n, 0, 63, st));
    void* temp = nullptr;
    RTC_CUDA(scratch.alloc((uint8_t**)&temp, tempBytes));
    RTC_CUDA(cub::DeviceRadixSort::SortPairs(temp, tempBytes, k, v, (int)n, 0, 63, st));
    A.keys = k.Current(); A.order = v.Current();
  }
  phase("sort");
  if (n > 1)
  {
    k_radix_tree<<<(n - 1 + kB - 1) / kB, kB, 0, st>>>(A.keys, (int)n, A.child, A.parent, A.range);
    ctx->kernelLaunches++;
  }
  k_fit_boxes<<<gridN, kB, 0, st>>>((int)n, A.order, A.triLo, A.triHi, A.child, A.parent, A.boxLo, A.boxHi, A.flags);
  ctx->kernelLaunches++;
  RTC_CUDA(cudaGetLastError());

  float hostBox[6]; uint32_t hostCounters[8];
  RTC_CUDA(cudaMemcpyAsync(hostBox, A.sceneBox, sizeof(hostBox), cudaMemcpyDeviceToHost, st));
  RTC_CUDA(cudaMemcpyAsync(hostCounters, d_counters, sizeof(hostCounters), cudaMemcpyDeviceToHost, st));
  RTC_CUDA(cudaStreamSynchronize(st));
  if (hostCounters[0]) RTC_FAIL("triangle index out of range");
  for (int k = 0; k < 3; ++k) { rec.lo[k] = hostBox[k]; rec.hi[k] = hostBox[3 + k]; }

  phase("tree+fit");
  // ---- SAH-binned top levels over a cut of the radix tree
  int rootNode = (n == 1) ? 0 : 0;            // binary root (for n == 1 the single leaf has id n-1 = 0)
  if (n > 8 * kLeafMax * 64)
  {
    // subtrees of at most n / kCutDivisor triangles form the cut whose top is rebuilt with binned SAH on the host
    uint32_t cutDivisor = 2048u;
    if (const char* e = std::getenv("RTC_GPU_CUT_DIVISOR")) { const long v = atol(e); if (64 <= v && v <= 32768) cutDivisor = (uint32_t)v; }
    int limit = (int)(n / cutDivisor); if (limit < (int)kLeafMax) limit = (int)kLeafMax;
    k_mark_cut<<<(unsigned)((numBinary + kB - 1) / kB), kB, 0, st>>>((int)n, limit, A.parent, A.range, d_cut, d_counters + 1, cutCapacity);
    RTC_CUDA(cudaMemcpyAsync(hostCounters, d_counters, sizeof(hostCounters), cudaMemcpyDeviceToHost, st));
    RTC_CUDA(cudaStreamSynchronize(st));
    const uint32_t cutCount = hostCounters[1];
    if (verbose) std::fprintf(stderr, "build_gas_gpu[%u tris] cut of %u subtrees (<= %d triangles each, capacity %u)\n", n, cutCount, limit, cutCapacity);
    if (cutCount >= 2 && cutCount <= cutCapacity)
    {
      k_gather_boxes<<<(cutCount + kB - 1) / kB, kB, 0, st>>>(d_cut, cutCount, A.boxLo, A.boxHi, d_cutBoxes);
      std::vector<PrimBox> cutBoxes(cutCount); std::vector<int> cutNodes(cutCount);
      RTC_CUDA(cudaMemcpyAsync(cutBoxes.data(), d_cutBoxes, cutCount * sizeof(PrimBox), cudaMemcpyDeviceToHost, st));
      RTC_CUDA(cudaMemcpyAsync(cutNodes.data(), d_cut, cutCount * sizeof(int), cudaMemcpyDeviceToHost, st));
      RTC_CUDA(cudaStreamSynchronize(st));
      std::vector<int2> topChildren; std::vector<PrimBox> topBoxes;
      build_binary_sah_host(cutBoxes.data(), cutCount, topChildren, topBoxes);
      // append: top node t -> id numBinary + t; its leaves are the cut's radix-tree nodes
      const size_t m = topChildren.size();
      std::vector<int2> hc(m), hr(m); std::vector<float4> hlo(m), hhi(m);
      for (size_t t = 0; t < m; ++t)
      {
        auto map = [&](int c) { return c < 0 ? cutNodes[(size_t)(~c)] : (int)(numBinary + (size_t)c); };
        hc[t] = make_int2(map(topChildren[t].x), map(topChildren[t].y));
        hr[t] = make_int2(0, (int)n - 1);        // only the size matters to the collapse: a top node is never a leaf cluster
        hlo[t] = make_float4(topBoxes[t].lo[0], topBoxes[t].lo[1], topBoxes[t].lo[2], 0.0f);
        hhi[t] = make_float4(topBoxes[t].hi[0], topBoxes[t].hi[1], topBoxes[t].hi[2], 0.0f);
      }
      RTC_CUDA(cudaMemcpyAsync(A.child + numBinary, hc.data(), m * sizeof(int2), cudaMemcpyHostToDevice, st));
      RTC_CUDA(cudaMemcpyAsync(A.range + numBinary, hr.data(), m * sizeof(int2), cudaMemcpyHostToDevice, st));
      RTC_CUDA(cudaMemcpyAsync(A.boxLo + numBinary, hlo.data(), m * sizeof(float4), cudaMemcpyHostToDevice, st));
      RTC_CUDA(cudaMemcpyAsync(A.boxHi + numBinary, hhi.data(), m * sizeof(float4), cudaMemcpyHostToDevice, st));
      RTC_CUDA(cudaStreamSynchronize(st));
      rootNode = (int)numBinary;
    }
  }

  phase("sah-top");
  // ---- collapse, level by level
  const size_t maxWide = (size_t)n + 16;          // every wide node but the root has a sibling group parent: < n nodes
  RTC_CUDA(cudaMalloc(&rec.d_nodes, maxWide * sizeof(Node8)));
  RTC_CUDA(cudaMalloc(&rec.d_tris, (size_t)n * 3u * sizeof(float4)));
  int2* d_work[2] = { nullptr, nullptr };
  RTC_CUDA(scratch.alloc(&d_work[0], maxWide)); RTC_CUDA(scratch.alloc(&d_work[1], maxWide));
  CollapseParams P;
  P.n = (int)n; P.child = A.child; P.range = A.range; P.boxLo = A.boxLo; P.boxHi = A.boxHi; P.order = A.order;
  P.verts = (const uint8_t*)(uintptr_t)rec.attributes; P.stride = rec.strideBytes; P.idx = (const uint32_t*)(uintptr_t)rec.indices;
  P.nodes = (Node8*)rec.d_nodes; P.tris = (float4*)rec.d_tris; P.nodeCount = d_counters + 2; P.triCount = d_counters + 3;
  // One triangle per leaf child.  A triangle test costs a warp ~2.7x what a node visit costs (it runs for the 3-4 lanes that
  // have just found a leaf, the node visit for 26-29), and the leaves of an LBVH over unstructured triangles are loose: single-
  // triangle leaves trade 18 triangle tests per ray for a few more node visits -- 1 M / 10 M-triangle soups, incoherent closest
  // hit: 1069 -> 1358 and 818 -> 1090 Mrays/s (leaves of <= 2: 1196 / 920).  RTC_GPU_LEAF_MAX=2|3 restores larger leaves.
  P.leafMax = 1;
  if (const char* e = std::getenv("RTC_GPU_LEAF_MAX")) { const int v = atoi(e); if (1 <= v && v <= (int)kLeafMax) P.leafMax = v; }
  const uint32_t one = 1u, zero = 0u;
  const int2 rootWork = make_int2(rootNode, 0);
  RTC_CUDA(cudaMemcpyAsync(d_counters + 2, &one, sizeof(uint32_t), cudaMemcpyHostToDevice, st));     // wide node 0 = root
  RTC_CUDA(cudaMemcpyAsync(d_counters + 3, &zero, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  RTC_CUDA(cudaMemcpyAsync(d_work[0], &rootWork, sizeof(int2), cudaMemcpyHostToDevice, st));
  uint32_t workCount = 1;
  int cur = 0, levels = 0;
  while (workCount)
  {
    ++levels;
    RTC_CUDA(cudaMemcpyAsync(d_counters + 4, &zero, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    k_collapse_level<<<(workCount + 127) / 128, 128, 0, st>>>(P, d_work[cur], workCount, d_work[cur ^ 1], d_counters + 4);
    ctx->kernelLaunches++;
    RTC_CUDA(cudaMemcpyAsync(&workCount, d_counters + 4, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    RTC_CUDA(cudaStreamSynchronize(st));
    cur ^= 1;
  }
  RTC_CUDA(cudaMemcpyAsync(hostCounters, d_counters, sizeof(hostCounters), cudaMemcpyDeviceToHost, st));
  RTC_CUDA(cudaStreamSynchronize(st));
  RTC_CUDA(cudaGetLastError());
  phase("collapse");
  if (verbose) std::fprintf(stderr, "build_gas_gpu[%u tris] %d collapse levels, %u wide nodes\n", n, levels, hostCounters[2]);
  if (hostCounters[3] != n) RTC_FAIL("GPU build lost triangles (internal error)");
  rec.numNodes = hostCounters[2];
  // the node array was sized for the worst case (one wide node per triangle): shrink it to what the collapse produced
  if ((size_t)rec.numNodes * 2 < maxWide)
  {
    void* exact = nullptr;
    RTC_CUDA(cudaMalloc(&exact, (size_t)rec.numNodes * sizeof(Node8)));
    cudaError_t e = cudaMemcpyAsync(exact, rec.d_nodes, (size_t)rec.numNodes * sizeof(Node8), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cudaFree(exact); return rtc_set_error(__FILE__, __LINE__, "shrink node array", (int)e, cudaGetErrorString(e)); }
    RTC_CUDA(cudaFree(rec.d_nodes));
    rec.d_nodes = exact;
  }
  return 0;
}

// The outputs are allocated half-way through the build: a failure after that point must not leak them.
int build_gas_gpu(rtc_context* ctx, GasRecord& rec)
{
  const int rc = build_gas_gpu_impl(ctx, rec);
  if (rc != 0)
  {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(rec.d_nodes); cudaFree(rec.d_tris);
    rec.d_nodes = nullptr; rec.d_tris = nullptr;
    cudaGetLastError();
  }
  return rc;
}

// ---- instance level: tight world bounds of every instance from its transformed vertices -------------------------
// The reference hands OptiX the instance transform and lets the driver bound the instance (Device.cpp:1427-1443,
// optixAccelBuild over OPTIX_BUILD_INPUT_TYPE_INSTANCES :1471-1482).  The first version here bounded the eight transformed
// corners of the GAS box, which is loose for rotated instances (a torus rotated by 45 degrees: +40 % per axis) and made
// rays enter instances they cannot hit; an instance entry costs about three node visits.  One CTA per instance now
// transforms every vertex of the instance's GAS and reduces min/max.
namespace {

struct InstanceBoundsIn
{
  float transform[12];
  const uint8_t* verts; uint32_t stride, numVerts;
  uint32_t pad;
};

__global__ void __launch_bounds__(kB)
k_instance_bounds(const InstanceBoundsIn* __restrict__ in, PrimBox* __restrict__ out)
{
  const InstanceBoundsIn& I = in[blockIdx.x];
  float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
  for (uint32_t v = threadIdx.x; v < I.numVerts; v += blockDim.x)
  {
    const float* p = reinterpret_cast<const float*>(I.verts + (size_t)v * I.stride);
    const float x = p[0], y = p[1], z = p[2];
#pragma unroll
    for (int r = 0; r < 3; ++r)
    {
      const float* m = I.transform + 4 * r;
      const float w = fmaf(m[0], x, fmaf(m[1], y, fmaf(m[2], z, m[3])));
      lo[r] = fminf(lo[r], w); hi[r] = fmaxf(hi[r], w);
    }
  }
  __shared__ float sLo[3][kB / 32], sHi[3][kB / 32];
#pragma unroll
  for (int r = 0; r < 3; ++r)
  {
    for (int off = 16; off; off >>= 1)
    {
      lo[r] = fminf(lo[r], __shfl_down_sync(0xffffffffu, lo[r], off));
      hi[r] = fmaxf(hi[r], __shfl_down_sync(0xffffffffu, hi[r], off));
    }
    if ((threadIdx.x & 31) == 0) { sLo[r][threadIdx.x >> 5] = lo[r]; sHi[r][threadIdx.x >> 5] = hi[r]; }
  }
  __syncthreads();
  if (threadIdx.x < 3)
  {
    float a = sLo[threadIdx.x][0], b = sHi[threadIdx.x][0];
    for (int w = 1; w < kB / 32; ++w) { a = fminf(a, sLo[threadIdx.x][w]); b = fmaxf(b, sHi[threadIdx.x][w]); }
    out[blockIdx.x].lo[threadIdx.x] = a; out[blockIdx.x].hi[threadIdx.x] = b;
  }
}

} // namespace

// boxes[i] = exact (unpadded) world bounds of the vertices of instance i; instances of an empty GAS get an inverted box
int instance_bounds_gpu(rtc_context* ctx, const rtc_instance_desc* instances, uint32_t numInstances, PrimBox* boxes)
{
  if (numInstances == 0) return 0;
  std::vector<InstanceBoundsIn> in(numInstances);
  for (uint32_t i = 0; i < numInstances; ++i)
  {
    const GasRecord& g = ctx->gas[instances[i].gas];
    for (int k = 0; k < 12; ++k) in[i].transform[k] = instances[i].transform[k];
    in[i].verts = reinterpret_cast<const uint8_t*>((uintptr_t)g.attributes); in[i].stride = g.strideBytes;
    in[i].numVerts = g.numTris ? g.numVerts : 0u; in[i].pad = 0u;
  }
  InstanceBoundsIn* d_in = nullptr; PrimBox* d_out = nullptr;
  RTC_CUDA(cudaMalloc(&d_in, sizeof(InstanceBoundsIn) * numInstances));
  if (cudaMalloc(&d_out, sizeof(PrimBox) * numInstances) != cudaSuccess) { cudaFree(d_in); RTC_FAIL("cudaMalloc failed"); }
  cudaError_t e = cudaMemcpyAsync(d_in, in.data(), sizeof(InstanceBoundsIn) * numInstances, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess)
  {
    k_instance_bounds<<<numInstances, kB, 0, ctx->stream>>>(d_in, d_out);
    ctx->kernelLaunches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(boxes, d_out, sizeof(PrimBox) * numInstances, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_in); cudaFree(d_out);
  if (e != cudaSuccess) return rtc_set_error(__FILE__, __LINE__, "instance_bounds_gpu", (int)e, cudaGetErrorString(e));
  return 0;
}
