// accel_host.cpp -- the host halves of the acceleration-structure builds, free of CUDA calls.
//
// rtc_gas_build (host SAH builder) and rtc_ias_build in rtc_api.cpp fetch their inputs from the device, call the functions
// below and upload what they return.  The same functions sit behind rtc_host_gas_build / rtc_host_ias_build
// (include/rtc_core.h), the host-only twin of the two builds: tools and the CPU tests build exactly the arrays a B200 would
// be handed -- wide nodes, leaf-ordered triangles, instance-level leaves, world->object matrices -- and the scalar oracle
// traverses them in the kernels' order of operations (oracle/wide_bvh.inc).  Nothing here renders or intersects anything.
#include "rtc_internal.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

// World->object matrix of an instance.  DEFINED here (OptiX computed it inside optixAccelBuild,
// Device.cpp:1478, read back by closesthit.cu:49-52): adjugate / determinant of the upper 3x3 in double,
// rounded once to float; translation -(Minv * t) in double.  The oracle states the same definition.
void invert_3x4(const float m[12], float out[12])
{
  const double a = m[0], b = m[1], c = m[2],  tx = m[3];
  const double d = m[4], e = m[5], f = m[6],  ty = m[7];
  const double g = m[8], h = m[9], i = m[10], tz = m[11];
  const double c00 = e * i - f * h, c01 = c * h - b * i, c02 = b * f - c * e;
  const double c10 = f * g - d * i, c11 = a * i - c * g, c12 = c * d - a * f;
  const double c20 = d * h - e * g, c21 = b * g - a * h, c22 = a * e - b * d;
  const double det = a * c00 + b * c10 + c * c20;
  const double r = 1.0 / det;
  const double i00 = c00 * r, i01 = c01 * r, i02 = c02 * r;
  const double i10 = c10 * r, i11 = c11 * r, i12 = c12 * r;
  const double i20 = c20 * r, i21 = c21 * r, i22 = c22 * r;
  out[0] = (float)i00; out[1] = (float)i01; out[2]  = (float)i02; out[3]  = (float)(-(i00 * tx + i01 * ty + i02 * tz));
  out[4] = (float)i10; out[5] = (float)i11; out[6]  = (float)i12; out[7]  = (float)(-(i10 * tx + i11 * ty + i12 * tz));
  out[8] = (float)i20; out[9] = (float)i21; out[10] = (float)i22; out[11] = (float)(-(i20 * tx + i21 * ty + i22 * tz));
}

// Host quality build of one geometry: triangle boxes, binned SAH, collapse, quantisation, triangles in leaf order
// (v0.xyz | primitive id bits, v1.xyz | 0, v2.xyz | 0).  Returns false when an index is out of range.
bool gas_assemble_host(const uint8_t* verts, uint32_t strideBytes, uint32_t numVerts, const uint32_t* idx, uint32_t numTris,
                       WideBvh& bvh, std::vector<float4>& tris)
{
  std::vector<PrimBox> boxes(numTris);
  auto vertex = [&](uint32_t i) { return reinterpret_cast<const float*>(verts + (size_t)i * strideBytes); };
  for (uint32_t t = 0; t < numTris; ++t)
  {
    PrimBox& b = boxes[t];
    for (int k = 0; k < 3; ++k) { b.lo[k] = std::numeric_limits<float>::infinity(); b.hi[k] = -b.lo[k]; }
    for (int c = 0; c < 3; ++c)
    {
      const uint32_t vi = idx[3u * t + c];
      if (vi >= numVerts) return false;
      const float* p = vertex(vi);
      for (int k = 0; k < 3; ++k) { b.lo[k] = std::fmin(b.lo[k], p[k]); b.hi[k] = std::fmax(b.hi[k], p[k]); }
    }
  }
  // triangles per leaf child of the host SAH build.  2: geometry scene 2526 -> 2545 Msamples/s, Cornell box 809 -> 820 against
  // leaves of up to 3; 1 is better only for the instanced scene (355 -> 368) and loses 2 % elsewhere.  RTC_HOST_LEAF_MAX overrides.
  uint32_t leafMax = 2;
  if (const char* e = getenv("RTC_HOST_LEAF_MAX")) { const int v = atoi(e); if (1 <= v && v <= 3) leafMax = (uint32_t)v; }
  build_wide_bvh_host(boxes.data(), numTris, bvh, leafMax);
  tris.resize((size_t)numTris * 3u);
  for (uint32_t s = 0; s < numTris; ++s)
  {
    const uint32_t prim = bvh.primOrder[s];
    for (int c = 0; c < 3; ++c)
    {
      const float* p = vertex(idx[3u * prim + c]);
      float w = 0.0f;
      if (c == 0) std::memcpy(&w, &prim, 4);
      tris[3u * (size_t)s + c] = make_float4(p[0], p[1], p[2], w);
    }
  }
  return true;
}

// Exact world bounds of the transformed vertices of one instance, in the arithmetic of k_instance_bounds (bvh_build_gpu.cu):
// one fmaf chain per row, min / max are order-independent, so the device reduction and this loop agree bit for bit.
void instance_bounds_host(const float transform[12], const uint8_t* verts, uint32_t strideBytes, uint32_t numVerts, PrimBox& out)
{
  float lo[3] = { std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max() };
  float hi[3] = { -lo[0], -lo[1], -lo[2] };
  for (uint32_t v = 0; v < numVerts; ++v)
  {
    const float* p = reinterpret_cast<const float*>(verts + (size_t)v * strideBytes);
    for (int r = 0; r < 3; ++r)
    {
      const float* m = transform + 4 * r;
      const float w = std::fmaf(m[0], p[0], std::fmaf(m[1], p[1], std::fmaf(m[2], p[2], m[3])));
      lo[r] = std::fmin(lo[r], w); hi[r] = std::fmax(hi[r], w);
    }
  }
  for (int r = 0; r < 3; ++r) { out.lo[r] = lo[r]; out.hi[r] = hi[r]; }
}

// The box the instance level is built over.  tight: b holds the exact bounds of the transformed vertices on entry and is padded
// -- the object-space ray is a ROUNDED transform of the world ray, so a hit found in object space may lie a few ulps outside
// the exact world-space box.  Otherwise b becomes the padded bounds of the eight transformed corners of the GAS box (the
// round-1 bounds, looser for rotated instances).  Instances of an empty GAS get a point box at the origin.
void instance_box_finish(const float transform[12], const float gasLo[3], const float gasHi[3], bool tight, bool emptyGas, PrimBox& b)
{
  const double ext = std::fabs((double)gasHi[0] - gasLo[0]) + std::fabs((double)gasHi[1] - gasLo[1]) + std::fabs((double)gasHi[2] - gasLo[2]);
  if (tight)
  {
    for (int r = 0; r < 3; ++r)
    {
      const float* m = &transform[4 * r];
      const double scale = std::fabs((double)m[0]) + std::fabs((double)m[1]) + std::fabs((double)m[2]);
      const double padLo = (std::fabs((double)b.lo[r]) + ext * scale) * 1.0e-5, padHi = (std::fabs((double)b.hi[r]) + ext * scale) * 1.0e-5;
      b.lo[r] = std::nextafterf((float)((double)b.lo[r] - padLo), -std::numeric_limits<float>::infinity());
      b.hi[r] = std::nextafterf((float)((double)b.hi[r] + padHi), std::numeric_limits<float>::infinity());
    }
  }
  else
  {
    for (int k = 0; k < 3; ++k) { b.lo[k] = std::numeric_limits<float>::infinity(); b.hi[k] = -b.lo[k]; }
    for (int corner = 0; corner < 8; ++corner)
    {
      const double x = (corner & 1) ? gasHi[0] : gasLo[0], y = (corner & 2) ? gasHi[1] : gasLo[1], z = (corner & 4) ? gasHi[2] : gasLo[2];
      for (int r = 0; r < 3; ++r)
      {
        const float* m = &transform[4 * r];
        const double wv = m[0] * x + m[1] * y + m[2] * z + m[3];
        const double scale = std::fabs((double)m[0]) + std::fabs((double)m[1]) + std::fabs((double)m[2]);
        const double pad = (std::fabs(wv) + ext * scale) * 1.0e-5;
        const float lo = std::nextafterf((float)(wv - pad), -std::numeric_limits<float>::infinity());
        const float hi = std::nextafterf((float)(wv + pad), std::numeric_limits<float>::infinity());
        b.lo[r] = std::fmin(b.lo[r], lo); b.hi[r] = std::fmax(b.hi[r], hi);
      }
    }
  }
  if (emptyGas) { for (int k = 0; k < 3; ++k) { b.lo[k] = 0.0f; b.hi[k] = 0.0f; } }
}

bool instance_bounds_tight()
{
  const char* e = getenv("RTC_INSTANCE_BOUNDS");
  return !(e && e[0] == 'b');
}

void tlas_build_host(const PrimBox* boxes, uint32_t numInstances, WideBvh& bvh)
{
  build_wide_bvh_host(boxes, numInstances, bvh, getenv("RTC_TLAS_LEAF") ? (uint32_t)atoi(getenv("RTC_TLAS_LEAF")) : 1u, true);
}

// ------------------------------------------------------------------------------------------------
// Host-only twin of rtc_gas_build(RTC_BUILD_HOST_SAH) + rtc_ias_build (include/rtc_core.h)
// ------------------------------------------------------------------------------------------------
struct rtc_host_accel
{
  bool isInstanceLevel = false;
  WideBvh bvh;
  std::vector<float4> tris;             // geometry level
  std::vector<uint8_t> positions;       // geometry level: vertex positions, 12 B apart (for the instance bounds)
  uint32_t numVerts = 0, numTris = 0;
  std::vector<float> inverses;          // instance level: 12 floats per instance
};

extern "C" {

int rtc_host_gas_build(const void* attributes, uint32_t strideBytes, uint32_t numVerts, const uint32_t* indices, uint32_t numTris,
                       rtc_host_accel** out)
{
  if (!out) RTC_FAIL("out is null");
  if (strideBytes < 12 || (strideBytes & 3u)) RTC_FAIL("vertex stride must be a multiple of 4 and at least 12");
  if ((numVerts && !attributes) || (numTris && !indices)) RTC_FAIL("null input array");
  rtc_host_accel* a = new rtc_host_accel();
  a->numVerts = numVerts; a->numTris = numTris;
  if (!gas_assemble_host(static_cast<const uint8_t*>(attributes), strideBytes, numVerts, indices, numTris, a->bvh, a->tris))
  {
    delete a;
    RTC_FAIL("triangle index out of range");
  }
  a->positions.resize((size_t)numVerts * 12u);
  for (uint32_t v = 0; v < numVerts; ++v) std::memcpy(&a->positions[(size_t)v * 12u], static_cast<const uint8_t*>(attributes) + (size_t)v * strideBytes, 12);
  *out = a;
  return 0;
}

int rtc_host_ias_build(const float* transforms, const rtc_host_accel* const* geometry, uint32_t numInstances, rtc_host_accel** out)
{
  if (!out) RTC_FAIL("out is null");
  if (numInstances && (!transforms || !geometry)) RTC_FAIL("null input array");
  rtc_host_accel* a = new rtc_host_accel();
  a->isInstanceLevel = true;
  a->inverses.resize((size_t)numInstances * 12u);
  std::vector<PrimBox> boxes(numInstances);
  const bool tight = instance_bounds_tight();
  for (uint32_t i = 0; i < numInstances; ++i)
  {
    const rtc_host_accel* g = geometry[i];
    if (!g || g->isInstanceLevel) { delete a; RTC_FAIL("bad geometry handle in instance"); }
    const float* m = transforms + 12u * (size_t)i;
    invert_3x4(m, &a->inverses[(size_t)i * 12u]);
    if (tight) instance_bounds_host(m, g->positions.data(), 12u, g->numTris ? g->numVerts : 0u, boxes[i]);
    instance_box_finish(m, g->bvh.lo, g->bvh.hi, tight, g->numTris == 0, boxes[i]);
  }
  tlas_build_host(boxes.data(), numInstances, a->bvh);
  *out = a;
  return 0;
}

int rtc_host_accel_info(const rtc_host_accel* accel, uint64_t* numNodes, uint64_t* numPrims, float bounds[6])
{
  if (!accel) RTC_FAIL("accel is null");
  if (numNodes) *numNodes = accel->bvh.nodes.size();
  if (numPrims) *numPrims = accel->bvh.primOrder.size();
  if (bounds) for (int k = 0; k < 3; ++k) { bounds[k] = accel->bvh.lo[k]; bounds[3 + k] = accel->bvh.hi[k]; }
  return 0;
}

int rtc_host_accel_export(const rtc_host_accel* accel, void* nodes, uint32_t* primOrder, float* tris, float* worldToObject)
{
  if (!accel) RTC_FAIL("accel is null");
  if (nodes) std::memcpy(nodes, accel->bvh.nodes.data(), accel->bvh.nodes.size() * sizeof(Node8));
  if (primOrder && !accel->bvh.primOrder.empty()) std::memcpy(primOrder, accel->bvh.primOrder.data(), accel->bvh.primOrder.size() * 4u);
  if (tris && !accel->tris.empty()) std::memcpy(tris, accel->tris.data(), accel->tris.size() * sizeof(float4));
  if (worldToObject && !accel->inverses.empty()) std::memcpy(worldToObject, accel->inverses.data(), accel->inverses.size() * sizeof(float));
  return 0;
}

void rtc_host_accel_destroy(rtc_host_accel* accel) { delete accel; }

} // extern "C"
