// rtc_api.cpp -- the C ABI of librtcore (include/rtc_core.h): contexts, memory, acceleration-structure
// builds, launches.  This file replaces the OptiX host calls of apps/rtigo3/src/Device.cpp (see the table
// at the top of rtc_core.h); there is no CPU rendering path here: every entry point that produces
// rays, hits or pixels ends in a kernel launch on the context's stream.
#include "rtc_internal.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <set>

namespace {

thread_local std::string g_lastError;

double now_ms()
{
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

SceneRecord* find_scene(rtc_context* ctx, uint64_t topObject)
{
  for (SceneRecord* s : ctx->scenes) if ((uint64_t)(uintptr_t)s->d_desc == topObject) return s;
  return nullptr;
}

void free_scene(SceneRecord* s)
{
  cudaFree(s->d_desc); cudaFree(s->d_tlasNodes); cudaFree(s->d_instances); cudaFree(s->d_tlasLeaves); cudaFree(s->d_o2w); cudaFree(s->d_geomInst); cudaFree(s->d_instFlags);
  delete s;
}

} // namespace

int rtc_set_error(const char* file, int line, const char* call, int code, const char* text)
{
  char buf[1024];
  std::snprintf(buf, sizeof(buf), "ERROR: %s(%d): %s (%d) %s", file, line, call, code, text ? text : "");
  g_lastError = buf;
  return code ? code : -1;
}

static int take_event(rtc_context* ctx, cudaEvent_t* e)
{
  if (!ctx->eventPool.empty()) { *e = ctx->eventPool.back(); ctx->eventPool.pop_back(); return 0; }
  RTC_CUDA(cudaEventCreate(e));
  return 0;
}

int profile_begin(rtc_context* ctx, int cls)
{
  if (!ctx->profiling) return 0;
  // bound the number of pending events: resolve (synchronise) every 4096 launches
  if (ctx->spans.size() >= 4096)
  {
    RTC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (rtc_context::ProfileSpan& sp : ctx->spans)
    {
      float ms = 0.0f;
      RTC_CUDA(cudaEventElapsedTime(&ms, sp.a, sp.b));
      ctx->profile.ms[sp.cls] += ms; ctx->profile.launches[sp.cls] += 1;
      ctx->eventPool.push_back(sp.a); ctx->eventPool.push_back(sp.b);
    }
    ctx->spans.clear();
  }
  rtc_context::ProfileSpan sp; sp.cls = cls;
  if (int rc = take_event(ctx, &sp.a)) return rc;
  if (int rc = take_event(ctx, &sp.b)) return rc;
  RTC_CUDA(cudaEventRecord(sp.a, ctx->stream));
  ctx->spans.push_back(sp);
  return 0;
}

int profile_end(rtc_context* ctx)
{
  if (!ctx->profiling) return 0;
  RTC_CUDA(cudaEventRecord(ctx->spans.back().b, ctx->stream));
  return 0;
}

extern "C" {

int rtc_version(void) { return 100; }

const char* rtc_last_error(void) { return g_lastError.c_str(); }

int rtc_context_destroy(rtc_context* ctx);

static int context_init(rtc_context* ctx, int deviceOrdinal)
{
  ctx->device = deviceOrdinal;
  cudaDeviceProp prop;
  RTC_CUDA(cudaGetDeviceProperties(&prop, deviceOrdinal));
  ctx->numSMs = prop.multiProcessorCount;
  // traversal driver: one ray per lane (default) or the per-warp ray pool; both are compiled, RTC_TRACE_DRIVER=lane|pool picks
  ctx->traceDriver = RTC_DRIVER_LANE;
  if (const char* e = getenv("RTC_PRIMARY_PACKETS")) ctx->primaryPackets = atoi(e) != 0;
  if (const char* e = getenv("RTC_TRACE_DRIVER")) ctx->traceDriver = (e[0] == 'p' || e[0] == '1') ? RTC_DRIVER_POOL : RTC_DRIVER_LANE;
  // schedule of the lane-owned driver's triangle tests: measured per context (auto) unless RTC_TRACE_SCHEDULE=group|onetri|twotri fixes it
  if (const char* e = getenv("RTC_TRACE_SCHEDULE"))
  {
    if (e[0] == 'g' || e[0] == '0') { ctx->traceSchedule = RTC_SCHEDULE_GROUP; ctx->tuner.state = ScheduleTuner::DONE; ctx->tuner.fixedByEnv = true; }
    else if (e[0] == 'o' || e[0] == '1') { ctx->traceSchedule = RTC_SCHEDULE_ONE_TRI; ctx->tuner.state = ScheduleTuner::DONE; ctx->tuner.fixedByEnv = true; }
    else if (e[0] == 't' || e[0] == '2') { ctx->traceSchedule = RTC_SCHEDULE_TWO_TRI; ctx->tuner.state = ScheduleTuner::DONE; ctx->tuner.fixedByEnv = true; }
  }
  RTC_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  {
    // keep freed build scratch in the device's default memory pool instead of returning it to the OS at every synchronise
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, deviceOrdinal) == cudaSuccess)
    {
      uint64_t threshold = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    cudaGetLastError();
  }
  RTC_CUDA(cudaEventCreate(&ctx->evA));
  RTC_CUDA(cudaEventCreate(&ctx->evB));
  RTC_CUDA(cudaEventCreate(&ctx->evTimerA));
  RTC_CUDA(cudaEventCreate(&ctx->evTimerB));
  RTC_CUDA(cudaMalloc(&ctx->d_stats, 8 * sizeof(uint64_t)));
  RTC_CUDA(cudaMemset(ctx->d_stats, 0, 8 * sizeof(uint64_t)));
  RTC_CUDA(cudaMalloc(&ctx->d_cursor, 4 * sizeof(uint32_t)));
  RTC_CUDA(cudaMalloc(&ctx->d_launchCounts, 3 * kTraceCountWords * sizeof(unsigned long long)));
  RTC_CUDA(cudaMemset(ctx->d_launchCounts, 0, 3 * kTraceCountWords * sizeof(unsigned long long)));
  return 0;
}

int rtc_context_create(int deviceOrdinal, rtc_context** out)
{
  if (!out) RTC_FAIL("out is null");
  *out = nullptr;
  int count = 0;
  RTC_CUDA(cudaGetDeviceCount(&count));
  if (deviceOrdinal < 0 || deviceOrdinal >= count) RTC_FAIL("device ordinal out of range (no CUDA device, no rendering: this core has no CPU path)");
  RTC_CUDA(cudaSetDevice(deviceOrdinal));
  rtc_context* ctx = new rtc_context();
  if (const int rc = context_init(ctx, deviceOrdinal))
  {
    const std::string why = rtc_last_error();      // the destroy below must not clobber the reason
    rtc_context_destroy(ctx);
    g_lastError = why;
    return rc;
  }
  *out = ctx;
  return 0;
}


int rtc_context_destroy(rtc_context* ctx)
{
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  release_cutout_graph(ctx);
  tuner_release(ctx);
  for (SceneRecord* s : ctx->scenes) free_scene(s);
  for (GasRecord& g : ctx->gas) { cudaFree(g.d_nodes); cudaFree(g.d_tris); }
  if (ctx->wf.base) cudaFree(ctx->wf.base);
  if (ctx->wf.cutBase) cudaFree(ctx->wf.cutBase);
  for (void* t : ctx->textures) cudaFree(t);
  cudaFree(ctx->d_stats);
  cudaFree(ctx->d_launchCounts);
  cudaFree(ctx->d_cursor);
  cudaFree(ctx->d_poolScratch);
  for (rtc_context::ProfileSpan& sp : ctx->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  for (cudaEvent_t e : ctx->eventPool) cudaEventDestroy(e);
  for (cudaEvent_t e : { ctx->evA, ctx->evB, ctx->evTimerA, ctx->evTimerB }) if (e) cudaEventDestroy(e);
  for (int k = 0; k < 6; ++k) { if (ctx->shadeStreams[k]) cudaStreamDestroy(ctx->shadeStreams[k]); if (ctx->shadeJoin[k]) cudaEventDestroy(ctx->shadeJoin[k]); }
  if (ctx->shadeFork) cudaEventDestroy(ctx->shadeFork);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  cudaGetLastError();
  delete ctx;
  return 0;
}

int rtc_synchronize(rtc_context* ctx)
{
  RTC_CUDA(cudaSetDevice(ctx->device));
  RTC_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

uint64_t rtc_context_stream(rtc_context* ctx) { return (uint64_t)(uintptr_t)ctx->stream; }

int rtc_device_count(int* count)
{
  if (!count) RTC_FAIL("count is null");
  *count = 0;
  RTC_CUDA(cudaGetDeviceCount(count));
  return 0;
}

int rtc_device_name(int deviceOrdinal, char* name, int length)
{
  if (!name || length <= 0) RTC_FAIL("bad name buffer");
  cudaDeviceProp prop;
  RTC_CUDA(cudaGetDeviceProperties(&prop, deviceOrdinal));
  std::snprintf(name, (size_t)length, "%s", prop.name);
  return 0;
}

int rtc_peer_can_access(int deviceOrdinal, int peerOrdinal, int* canAccess)
{
  if (!canAccess) RTC_FAIL("canAccess is null");
  if (deviceOrdinal == peerOrdinal) { *canAccess = 1; return 0; }
  RTC_CUDA(cudaDeviceCanAccessPeer(canAccess, deviceOrdinal, peerOrdinal));
  return 0;
}

int rtc_peer_enable(rtc_context* ctx, rtc_context* peer)
{
  if (ctx->device == peer->device) return 0;
  RTC_CUDA(cudaSetDevice(ctx->device));
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
  RTC_CUDA(e);
  return 0;
}

int rtc_peer_disable(rtc_context* ctx, rtc_context* peer)
{
  if (ctx->device == peer->device) return 0;
  RTC_CUDA(cudaSetDevice(ctx->device));
  const cudaError_t e = cudaDeviceDisablePeerAccess(peer->device);
  if (e == cudaErrorPeerAccessNotEnabled) { cudaGetLastError(); return 0; }
  RTC_CUDA(e);
  return 0;
}

int rtc_memcpy_peer(rtc_context* dstCtx, uint64_t dst, rtc_context* srcCtx, uint64_t src, uint64_t bytes)
{
  RTC_CUDA(cudaSetDevice(srcCtx->device));
  RTC_CUDA(cudaStreamSynchronize(srcCtx->stream));
  RTC_CUDA(cudaSetDevice(dstCtx->device));
  if (bytes == 0) return 0;
  if (dstCtx->device == srcCtx->device)
    RTC_CUDA(cudaMemcpyAsync((void*)(uintptr_t)dst, (const void*)(uintptr_t)src, bytes, cudaMemcpyDeviceToDevice, dstCtx->stream));
  else
    RTC_CUDA(cudaMemcpyPeerAsync((void*)(uintptr_t)dst, dstCtx->device, (const void*)(uintptr_t)src, srcCtx->device, bytes, dstCtx->stream));
  return 0;
}

int rtc_malloc(rtc_context* ctx, uint64_t bytes, uint64_t* dptr)
{
  RTC_CUDA(cudaSetDevice(ctx->device));
  void* p = nullptr;
  RTC_CUDA(cudaMalloc(&p, bytes ? bytes : 1));
  *dptr = (uint64_t)(uintptr_t)p;
  return 0;
}

int rtc_free(rtc_context* ctx, uint64_t dptr)
{
  RTC_CUDA(cudaSetDevice(ctx->device));
  RTC_CUDA(cudaStreamSynchronize(ctx->stream));
  RTC_CUDA(cudaFree((void*)(uintptr_t)dptr));
  return 0;
}

int rtc_upload(rtc_context* ctx, uint64_t dst, const void* src, uint64_t bytes)
{
  RTC_CUDA(cudaSetDevice(ctx->device));
  if (bytes) RTC_CUDA(cudaMemcpyAsync((void*)(uintptr_t)dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

int rtc_download(rtc_context* ctx, void* dst, uint64_t src, uint64_t bytes)
{
  RTC_CUDA(cudaSetDevice(ctx->device));
  if (bytes) RTC_CUDA(cudaMemcpyAsync(dst, (const void*)(uintptr_t)src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return 0;
}

int rtc_memset(rtc_context* ctx, uint64_t dst, int value, uint64_t bytes)
{
  RTC_CUDA(cudaSetDevice(ctx->device));
  if (bytes) RTC_CUDA(cudaMemsetAsync((void*)(uintptr_t)dst, value, bytes, ctx->stream));
  return 0;
}

int rtc_host_alloc(rtc_context* ctx, uint64_t bytes, void** ptr)
{
  RTC_CUDA(cudaSetDevice(ctx->device));
  RTC_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped));
  return 0;
}

int rtc_host_free(rtc_context* ctx, void* ptr)
{
  RTC_CUDA(cudaSetDevice(ctx->device));
  RTC_CUDA(cudaFreeHost(ptr));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Geometry acceleration structure (replaces optixAccelBuild over OPTIX_BUILD_INPUT_TYPE_TRIANGLES,
// Device.cpp:1362-1405: vertex stride sizeof(TriangleAttributes), uint3 indices, flags NONE).
// ------------------------------------------------------------------------------------------------
int rtc_gas_build(rtc_context* ctx, uint64_t attributes, uint32_t strideBytes, uint32_t numVerts,
                  uint64_t indices, uint32_t numTris, uint32_t buildFlags, uint32_t* gas)
{
  if (!gas) RTC_FAIL("gas is null");
  if (strideBytes < 12 || (strideBytes & 3u)) RTC_FAIL("vertex stride must be a multiple of 4 and at least 12");
  RTC_CUDA(cudaSetDevice(ctx->device));
  const double t0 = now_ms();
  GasRecord rec;
  rec.attributes = attributes; rec.indices = indices; rec.strideBytes = strideBytes; rec.numVerts = numVerts; rec.numTris = numTris;

  int builder = (int)buildFlags;
  if (builder == RTC_BUILD_DEFAULT)
  {
    builder = (numTris > kGpuBuildThreshold) ? RTC_BUILD_GPU_LBVH : RTC_BUILD_HOST_SAH;
    // RTC_FORCE_BUILDER=gpu|host overrides the size rule (used by the tests to run whole scenes through either builder)
    if (const char* force = std::getenv("RTC_FORCE_BUILDER"))
    {
      if (std::strcmp(force, "gpu") == 0) builder = RTC_BUILD_GPU_LBVH;
      else if (std::strcmp(force, "host") == 0) builder = RTC_BUILD_HOST_SAH;
    }
  }
  if (builder != RTC_BUILD_GPU_LBVH && builder != RTC_BUILD_HOST_SAH) RTC_FAIL("bad build flags");
  rec.builder = builder;

  if (builder == RTC_BUILD_GPU_LBVH && numTris > 0)
  {
    if (int rc = build_gas_gpu(ctx, rec)) return rc;
  }
  else
  {
    // host quality build: fetch positions and indices, bin-SAH, collapse, quantise, upload
    std::vector<uint8_t> verts((size_t)numVerts * strideBytes);
    std::vector<uint32_t> idx((size_t)numTris * 3u);
    RTC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!verts.empty()) RTC_CUDA(cudaMemcpy(verts.data(), (const void*)(uintptr_t)attributes, verts.size(), cudaMemcpyDeviceToHost));
    if (!idx.empty()) RTC_CUDA(cudaMemcpy(idx.data(), (const void*)(uintptr_t)indices, idx.size() * 4u, cudaMemcpyDeviceToHost));
    WideBvh bvh;
    std::vector<float4> tris;
    if (!gas_assemble_host(verts.data(), strideBytes, numVerts, idx.data(), numTris, bvh, tris)) RTC_FAIL("triangle index out of range");
    rec.numNodes = (uint32_t)bvh.nodes.size();
    for (int k = 0; k < 3; ++k) { rec.lo[k] = bvh.lo[k]; rec.hi[k] = bvh.hi[k]; }
    RTC_CUDA(cudaMalloc(&rec.d_nodes, bvh.nodes.size() * sizeof(Node8)));
    RTC_CUDA(cudaMemcpy(rec.d_nodes, bvh.nodes.data(), bvh.nodes.size() * sizeof(Node8), cudaMemcpyHostToDevice));
    RTC_CUDA(cudaMalloc(&rec.d_tris, tris.empty() ? 16 : tris.size() * sizeof(float4)));
    if (!tris.empty()) RTC_CUDA(cudaMemcpy(rec.d_tris, tris.data(), tris.size() * sizeof(float4), cudaMemcpyHostToDevice));
  }
  rec.buildMs = now_ms() - t0;
  ctx->gas.push_back(rec);
  *gas = (uint32_t)(ctx->gas.size() - 1);
  return 0;
}

int rtc_gas_destroy(rtc_context* ctx, uint32_t gas)
{
  if (gas >= ctx->gas.size()) RTC_FAIL("bad gas handle");
  RTC_CUDA(cudaSetDevice(ctx->device));
  RTC_CUDA(cudaStreamSynchronize(ctx->stream));
  GasRecord& g = ctx->gas[gas];
  cudaFree(g.d_nodes); cudaFree(g.d_tris);
  g.d_nodes = nullptr; g.d_tris = nullptr; g.numNodes = 0; g.numTris = 0;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Instance level (replaces createInstance + optixAccelBuild INSTANCES + createHitGroupRecords,
// Device.cpp:1427-1532).  Single-level instancing, as the reference's pipeline (Device.cpp:560).
// ------------------------------------------------------------------------------------------------
int rtc_ias_build(rtc_context* ctx, const rtc_instance_desc* instances, uint32_t numInstances, uint64_t* topObject)
{
  if (!topObject) RTC_FAIL("topObject is null");
  RTC_CUDA(cudaSetDevice(ctx->device));
  const double t0 = now_ms();
  std::vector<PrimBox> boxes(numInstances);
  std::vector<float4> inst((size_t)numInstances * 4u), o2w((size_t)numInstances * 3u);
  std::vector<rt_GeometryInstanceData> gi(numInstances);
  SceneRecord* rec = new SceneRecord();
  rec->inverses.resize((size_t)numInstances * 12u);
  rec->instGas.resize(numInstances);
  std::set<uint32_t> distinct;
  for (uint32_t i = 0; i < numInstances; ++i)
  {
    const rtc_instance_desc& d = instances[i];
    if (d.instanceId != i) { delete rec; RTC_FAIL("instanceId must equal the array position"); }
    if (d.gas >= ctx->gas.size() || ctx->gas[d.gas].d_nodes == nullptr) { delete rec; RTC_FAIL("bad gas handle in instance"); }
    const GasRecord& g = ctx->gas[d.gas];
    distinct.insert(d.gas);
    rec->instGas[i] = d.gas;
    float* inv = &rec->inverses[(size_t)i * 12u];
    invert_3x4(d.transform, inv);
    for (int r = 0; r < 3; ++r)
    {
      inst[4u * i + r] = make_float4(inv[4 * r], inv[4 * r + 1], inv[4 * r + 2], inv[4 * r + 3]);
      o2w[3u * i + r] = make_float4(d.transform[4 * r], d.transform[4 * r + 1], d.transform[4 * r + 2], d.transform[4 * r + 3]);
    }
    const uint64_t np = (uint64_t)(uintptr_t)g.d_nodes, tp = (uint64_t)(uintptr_t)g.d_tris;
    const uint32_t w[4] = { (uint32_t)np, (uint32_t)(np >> 32), (uint32_t)tp, (uint32_t)(tp >> 32) };
    std::memcpy(&inst[4u * i + 3], w, 16);
    gi[i].attributes = g.attributes; gi[i].indices = g.indices; gi[i].materialIndex = d.materialIndex; gi[i].lightIndex = d.lightIndex;

  }
  // World bounds per instance: exact bounds of the transformed vertices (device kernel), padded by instance_box_finish
  // (accel_host.cpp).  RTC_INSTANCE_BOUNDS=box falls back to the eight transformed corners of the GAS box.
  const bool tight = instance_bounds_tight();
  if (tight) { if (int rc = instance_bounds_gpu(ctx, instances, numInstances, boxes.data())) { delete rec; return rc; } }
  for (uint32_t i = 0; i < numInstances; ++i)
  {
    const GasRecord& g = ctx->gas[instances[i].gas];
    instance_box_finish(instances[i].transform, g.lo, g.hi, tight, g.numTris == 0, boxes[i]);
  }
  WideBvh bvh;
  tlas_build_host(boxes.data(), numInstances, bvh);

  auto upload = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, bytes ? bytes : 16);
    if (e != cudaSuccess) return e;
    return bytes ? cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) : cudaSuccess;
  };
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = upload(&rec->d_tlasNodes, bvh.nodes.data(), bvh.nodes.size() * sizeof(Node8));
  if (e == cudaSuccess) e = upload(&rec->d_tlasLeaves, bvh.primOrder.data(), bvh.primOrder.size() * 4u);
  if (e == cudaSuccess) e = upload(&rec->d_instances, inst.data(), inst.size() * sizeof(float4));
  if (e == cudaSuccess) e = upload(&rec->d_o2w, o2w.data(), o2w.size() * sizeof(float4));
  if (e == cudaSuccess) e = upload(&rec->d_geomInst, gi.data(), gi.size() * sizeof(rt_GeometryInstanceData));
  rec->instFlags.assign(numInstances, 0u);
  if (e == cudaSuccess) e = upload(&rec->d_instFlags, rec->instFlags.data(), rec->instFlags.size() * 4u);
  rec->desc.instFlags = (const uint32_t*)rec->d_instFlags;
  rec->desc.tlasNodes = (const uint4*)rec->d_tlasNodes;
  rec->desc.tlasLeaves = (const uint32_t*)rec->d_tlasLeaves;
  rec->desc.instances = (const float4*)rec->d_instances;
  rec->desc.objectToWorld = (const float4*)rec->d_o2w;
  rec->desc.geomInst = (const rt_GeometryInstanceData*)rec->d_geomInst;
  rec->desc.numInstances = numInstances;
  rec->desc.numTlasNodes = (uint32_t)bvh.nodes.size();
  rec->desc.numTlasLeaves = (uint32_t)bvh.primOrder.size();
  if (e == cudaSuccess) e = upload((void**)&rec->d_desc, &rec->desc, sizeof(SceneDesc));
  if (e != cudaSuccess) { free_scene(rec); return rtc_set_error(__FILE__, __LINE__, "rtc_ias_build upload", (int)e, cudaGetErrorString(e)); }
  rec->totalNodes = bvh.nodes.size(); rec->totalTris = 0; rec->gasBuildMs = 0.0; rec->numGas = (uint32_t)distinct.size();
  for (uint32_t g : distinct) { rec->totalNodes += ctx->gas[g].numNodes; rec->totalTris += ctx->gas[g].numTris; rec->gasBuildMs += ctx->gas[g].buildMs; }
  rec->iasBuildMs = now_ms() - t0;
  ctx->scenes.push_back(rec);
  *topObject = (uint64_t)(uintptr_t)rec->d_desc;
  return 0;
}

int rtc_scene_info_get(rtc_context* ctx, uint64_t topObject, rtc_scene_info* info)
{
  SceneRecord* s = find_scene(ctx, topObject);
  if (!s) RTC_FAIL("unknown topObject");
  std::memset(info, 0, sizeof(*info));
  info->numNodes = s->totalNodes; info->numTris = s->totalTris; info->numInstances = s->desc.numInstances;
  info->numTlasNodes = s->desc.numTlasNodes; info->numTlasLeaves = s->desc.numTlasLeaves; info->numGas = s->numGas; info->gasBuildMs = s->gasBuildMs; info->iasBuildMs = s->iasBuildMs;
  return 0;
}

int rtc_scene_destroy(rtc_context* ctx, uint64_t topObject)
{
  RTC_CUDA(cudaSetDevice(ctx->device));
  for (size_t i = 0; i < ctx->scenes.size(); ++i)
    if ((uint64_t)(uintptr_t)ctx->scenes[i]->d_desc == topObject)
    {
      RTC_CUDA(cudaStreamSynchronize(ctx->stream));
      free_scene(ctx->scenes[i]);
      ctx->scenes.erase(ctx->scenes.begin() + (long)i);
      return 0;
    }
  RTC_FAIL("unknown topObject");
}

// Per-instance hit-record selection (Device.cpp:1503-1513; updateMaterial :1141-1160 rewrites the SBT headers).
int rtc_scene_set_instance_flags(rtc_context* ctx, uint64_t topObject, uint32_t first, uint32_t count, const uint32_t* flags)
{
  SceneRecord* s = find_scene(ctx, topObject);
  if (!s) RTC_FAIL("unknown topObject");
  if ((uint64_t)first + count > s->desc.numInstances) RTC_FAIL("instance range out of bounds");
  if (count == 0) return 0;
  if (!flags) RTC_FAIL("flags is null");
  RTC_CUDA(cudaSetDevice(ctx->device));
  for (uint32_t i = 0; i < count; ++i) s->instFlags[first + i] = flags[i];
  s->numCutout = 0;
  for (uint32_t f : s->instFlags) if (f & RTC_INSTANCE_CUTOUT) s->numCutout++;
  // stream-ordered behind the launches already enqueued; `flags` may be reused by the caller at once (pageable source: staged copy)
  RTC_CUDA(cudaMemcpyAsync((uint32_t*)s->d_instFlags + first, &s->instFlags[first], (size_t)count * 4u, cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

int rtc_scene_set_albedo_textures(rtc_context* ctx, uint64_t topObject, int enable)
{
  SceneRecord* s = find_scene(ctx, topObject);
  if (!s) RTC_FAIL("unknown topObject");
  s->albedoTextures = enable != 0;
  return 0;
}

// Material textures: 16-byte header {width, height, 0, 0} + RGBA32F texels in one allocation; the handle is its address.
int rtc_texture_create(rtc_context* ctx, uint32_t width, uint32_t height, const float* rgba, uint64_t* handle)
{
  if (!handle) RTC_FAIL("handle is null");
  if (width == 0 || height == 0 || !rgba) RTC_FAIL("empty texture");
  RTC_CUDA(cudaSetDevice(ctx->device));
  const size_t texelBytes = (size_t)width * height * 16u;
  void* d = nullptr;
  RTC_CUDA(cudaMalloc(&d, 16u + texelBytes));
  const uint32_t header[4] = { width, height, 0u, 0u };
  cudaError_t e = cudaMemcpy(d, header, 16u, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy((char*)d + 16, rgba, texelBytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(d); return rtc_set_error(__FILE__, __LINE__, "rtc_texture_create upload", (int)e, cudaGetErrorString(e)); }
  ctx->textures.push_back(d);
  *handle = (uint64_t)(uintptr_t)d;
  return 0;
}

int rtc_texture_destroy(rtc_context* ctx, uint64_t handle)
{
  RTC_CUDA(cudaSetDevice(ctx->device));
  for (size_t i = 0; i < ctx->textures.size(); ++i)
    if ((uint64_t)(uintptr_t)ctx->textures[i] == handle)
    {
      RTC_CUDA(cudaStreamSynchronize(ctx->stream));
      RTC_CUDA(cudaFree(ctx->textures[i]));
      ctx->textures.erase(ctx->textures.begin() + (long)i);
      return 0;
    }
  RTC_FAIL("unknown texture handle");
}

int rtc_scene_export(rtc_context* ctx, uint64_t topObject, void* tlasNodes, uint32_t* tlasLeaves, float* worldToObject, uint32_t* instanceGas)
{
  SceneRecord* s = find_scene(ctx, topObject);
  if (!s) RTC_FAIL("unknown topObject");
  RTC_CUDA(cudaSetDevice(ctx->device));
  RTC_CUDA(cudaStreamSynchronize(ctx->stream));
  if (tlasNodes) RTC_CUDA(cudaMemcpy(tlasNodes, s->d_tlasNodes, (size_t)s->desc.numTlasNodes * sizeof(Node8), cudaMemcpyDeviceToHost));
  if (tlasLeaves && s->desc.numTlasLeaves) RTC_CUDA(cudaMemcpy(tlasLeaves, s->d_tlasLeaves, (size_t)s->desc.numTlasLeaves * 4u, cudaMemcpyDeviceToHost));
  if (worldToObject) std::memcpy(worldToObject, s->inverses.data(), s->inverses.size() * sizeof(float));
  if (instanceGas) std::memcpy(instanceGas, s->instGas.data(), s->instGas.size() * sizeof(uint32_t));
  return 0;
}

int rtc_gas_info(rtc_context* ctx, uint32_t gas, uint64_t* numNodes, uint64_t* numTris)
{
  if (gas >= ctx->gas.size() || ctx->gas[gas].d_nodes == nullptr) RTC_FAIL("bad gas handle");
  if (numNodes) *numNodes = ctx->gas[gas].numNodes;
  if (numTris) *numTris = ctx->gas[gas].numTris;
  return 0;
}

int rtc_gas_export(rtc_context* ctx, uint32_t gas, void* nodes, float* tris)
{
  if (gas >= ctx->gas.size() || ctx->gas[gas].d_nodes == nullptr) RTC_FAIL("bad gas handle");
  const GasRecord& g = ctx->gas[gas];
  RTC_CUDA(cudaSetDevice(ctx->device));
  RTC_CUDA(cudaStreamSynchronize(ctx->stream));
  if (nodes) RTC_CUDA(cudaMemcpy(nodes, g.d_nodes, (size_t)g.numNodes * sizeof(Node8), cudaMemcpyDeviceToHost));
  if (tris && g.numTris) RTC_CUDA(cudaMemcpy(tris, g.d_tris, (size_t)g.numTris * 12u * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

int rtc_instance_inverse(rtc_context* ctx, uint64_t topObject, uint32_t instance, float out[12])
{
  SceneRecord* s = find_scene(ctx, topObject);
  if (!s) RTC_FAIL("unknown topObject");
  if (instance >= s->desc.numInstances) RTC_FAIL("instance out of range");
  std::memcpy(out, &s->inverses[(size_t)instance * 12u], 48);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Launches
// ------------------------------------------------------------------------------------------------
int rtc_launch(rtc_context* ctx, const rt_SystemData* sys, uint32_t launchWidth, uint32_t launchHeight,
               int raygen, int miss, int iterationFirst, int iterationCount)
{
  if (!sys) RTC_FAIL("sys is null");
  if (raygen != RTC_RAYGEN_FULL_FRAME && raygen != RTC_RAYGEN_LOCAL_COPY) RTC_FAIL("bad raygen selector");
  if (miss < RT_MISS_NULL || miss > RT_MISS_SPHERE) RTC_FAIL("bad miss selector");
  if (miss == RT_MISS_SPHERE && (sys->envTexture == 0 || sys->envCDF_U == 0 || sys->envCDF_V == 0)) RTC_FAIL("miss 2 needs envTexture/envCDF_U/envCDF_V");
  RTC_CUDA(cudaSetDevice(ctx->device));
  return launch_wavefront(ctx, *sys, launchWidth, launchHeight, raygen, miss, iterationFirst, iterationCount, iterationFirst, false);
}

int rtc_launch_ex(rtc_context* ctx, const rt_SystemData* sys, uint32_t launchWidth, uint32_t launchHeight,
                  int raygen, int miss, int iterationFirst, int iterationCount, int accumulationFirst, int countWork)
{
  if (!sys) RTC_FAIL("sys is null");
  if (raygen != RTC_RAYGEN_FULL_FRAME && raygen != RTC_RAYGEN_LOCAL_COPY) RTC_FAIL("bad raygen selector");
  if (miss < RT_MISS_NULL || miss > RT_MISS_SPHERE) RTC_FAIL("bad miss selector");
  if (miss == RT_MISS_SPHERE && (sys->envTexture == 0 || sys->envCDF_U == 0 || sys->envCDF_V == 0)) RTC_FAIL("miss 2 needs envTexture/envCDF_U/envCDF_V");
  if (accumulationFirst < 0) RTC_FAIL("accumulationFirst < 0");
  RTC_CUDA(cudaSetDevice(ctx->device));
  return launch_wavefront(ctx, *sys, launchWidth, launchHeight, raygen, miss, iterationFirst, iterationCount, accumulationFirst, countWork != 0);
}

int rtc_trace_schedule_get(rtc_context* ctx, rtc_trace_schedule* out)
{
  if (!out) RTC_FAIL("out is null");
  RTC_CUDA(cudaSetDevice(ctx->device));
  tuner_finish(ctx);        // a pending decision is taken now (waits for the last timed batch)
  const ScheduleTuner& t = ctx->tuner;
  out->schedule = ctx->traceSchedule;
  out->decided = t.state == ScheduleTuner::DONE ? 1 : 0;
  out->measured = (t.state == ScheduleTuner::DONE && t.ms[1] > 0.0f) ? 1 : 0;
  out->pathsPerBatch = t.paths;
  out->groupMs[0] = t.ms[0]; out->oneTriMs = t.ms[1]; out->twoTriMs = t.ms[2]; out->groupMs[1] = t.ms[3];
  return 0;
}

int rtc_trace_schedule_set(rtc_context* ctx, int schedule)
{
  RTC_CUDA(cudaSetDevice(ctx->device));
  if (schedule == RTC_SCHEDULE_GROUP || schedule == RTC_SCHEDULE_ONE_TRI || schedule == RTC_SCHEDULE_TWO_TRI)
  {
    ctx->traceSchedule = schedule;
    ctx->tuner.state = ScheduleTuner::DONE;
    return 0;
  }
  if (schedule != -1) RTC_FAIL("schedule must be one of RTC_SCHEDULE_* or -1 (measure again)");
  ctx->traceSchedule = RTC_SCHEDULE_GROUP;
  ctx->tuner.state = ScheduleTuner::WARMUP;
  ctx->tuner.slot = 0; ctx->tuner.restarts = 0; ctx->tuner.paths = 0;
  for (float& ms : ctx->tuner.ms) ms = 0.0f;
  return 0;
}

int rtc_launch_counts_get(rtc_context* ctx, rtc_trace_counts out[2])
{
  if (!out) RTC_FAIL("out is null");
  RTC_CUDA(cudaSetDevice(ctx->device));
  unsigned long long h[2 * kTraceCountWords];
  RTC_CUDA(cudaMemcpyAsync(h, ctx->d_launchCounts, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  RTC_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < 2; ++k)
  {
    const unsigned long long* c = h + kTraceCountWords * k;
    out[k].nodes = c[0]; out[k].tris = c[1]; out[k].instances = c[2]; out[k].rays = c[3];
  }
  return 0;
}

int rtc_launch_pass_stats_get(rtc_context* ctx, rtc_pass_stats out[2])
{
  if (!out) RTC_FAIL("out is null");
  RTC_CUDA(cudaSetDevice(ctx->device));
  unsigned long long h[2 * kTraceCountWords];
  RTC_CUDA(cudaMemcpyAsync(h, ctx->d_launchCounts, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  RTC_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < 2; ++k)
    for (int j = 0; j < 4; ++j) { out[k].passes[j] = h[kTraceCountWords * k + 4 + j]; out[k].lanes[j] = h[kTraceCountWords * k + 8 + j]; }
  return 0;
}

int rtc_launch_counts_reset(rtc_context* ctx)
{
  RTC_CUDA(cudaSetDevice(ctx->device));
  RTC_CUDA(cudaMemsetAsync(ctx->d_launchCounts, 0, 2 * kTraceCountWords * sizeof(unsigned long long), ctx->stream));
  return 0;
}

int rtc_timer_start(rtc_context* ctx)
{
  RTC_CUDA(cudaSetDevice(ctx->device));
  RTC_CUDA(cudaEventRecord(ctx->evTimerA, ctx->stream));
  return 0;
}

int rtc_timer_stop(rtc_context* ctx, float* milliseconds)
{
  if (!milliseconds) RTC_FAIL("milliseconds is null");
  RTC_CUDA(cudaSetDevice(ctx->device));
  RTC_CUDA(cudaEventRecord(ctx->evTimerB, ctx->stream));
  RTC_CUDA(cudaEventSynchronize(ctx->evTimerB));
  RTC_CUDA(cudaEventElapsedTime(milliseconds, ctx->evTimerA, ctx->evTimerB));
  return 0;
}

static int profile_resolve(rtc_context* ctx)
{
  RTC_CUDA(cudaStreamSynchronize(ctx->stream));
  for (rtc_context::ProfileSpan& sp : ctx->spans)
  {
    float ms = 0.0f;
    RTC_CUDA(cudaEventElapsedTime(&ms, sp.a, sp.b));
    ctx->profile.ms[sp.cls] += ms;
    ctx->profile.launches[sp.cls] += 1;
    ctx->eventPool.push_back(sp.a); ctx->eventPool.push_back(sp.b);
  }
  ctx->spans.clear();
  return 0;
}

int rtc_profile_enable(rtc_context* ctx, int enable)
{
  RTC_CUDA(cudaSetDevice(ctx->device));
  if (int rc = profile_resolve(ctx)) return rc;
  if (enable) std::memset(&ctx->profile, 0, sizeof(ctx->profile));
  ctx->profiling = enable != 0;
  return 0;
}

int rtc_profile_get(rtc_context* ctx, rtc_profile* out)
{
  if (!out) RTC_FAIL("out is null");
  RTC_CUDA(cudaSetDevice(ctx->device));
  if (int rc = profile_resolve(ctx)) return rc;
  *out = ctx->profile;
  return 0;
}

int rtc_trace_closest(rtc_context* ctx, uint64_t topObject, uint64_t rays, uint64_t numRays, uint64_t hits)
{
  SceneRecord* s = find_scene(ctx, topObject);
  if (!s) RTC_FAIL("unknown topObject");
  RTC_CUDA(cudaSetDevice(ctx->device));
  RTC_CUDA(cudaEventRecord(ctx->evA, ctx->stream));
  if (int rc = launch_trace_closest(ctx, &s->desc, (const rtc_ray*)(uintptr_t)rays, numRays, (rtc_hit*)(uintptr_t)hits)) return rc;
  RTC_CUDA(cudaEventRecord(ctx->evB, ctx->stream));
  ctx->traceTimed = true;
  return 0;
}

int rtc_trace_any(rtc_context* ctx, uint64_t topObject, uint64_t rays, uint64_t numRays, uint64_t occluded)
{
  SceneRecord* s = find_scene(ctx, topObject);
  if (!s) RTC_FAIL("unknown topObject");
  RTC_CUDA(cudaSetDevice(ctx->device));
  RTC_CUDA(cudaEventRecord(ctx->evA, ctx->stream));
  if (int rc = launch_trace_any(ctx, &s->desc, (const rtc_ray*)(uintptr_t)rays, numRays, (uint32_t*)(uintptr_t)occluded)) return rc;
  RTC_CUDA(cudaEventRecord(ctx->evB, ctx->stream));
  ctx->traceTimed = true;
  return 0;
}

int rtc_trace_count(rtc_context* ctx, uint64_t topObject, uint64_t rays, uint64_t numRays, int anyHit, rtc_trace_counts* out)
{
  SceneRecord* s = find_scene(ctx, topObject);
  if (!s) RTC_FAIL("unknown topObject");
  if (!out) RTC_FAIL("out is null");
  RTC_CUDA(cudaSetDevice(ctx->device));
  unsigned long long* d = ctx->d_launchCounts + 2 * kTraceCountWords;
  RTC_CUDA(cudaMemsetAsync(d, 0, kTraceCountWords * sizeof(uint64_t), ctx->stream));
  if (int rc = launch_trace_count(ctx, &s->desc, (const rtc_ray*)(uintptr_t)rays, numRays, anyHit, d)) return rc;
  uint64_t h[4];
  RTC_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  RTC_CUDA(cudaStreamSynchronize(ctx->stream));
  out->nodes = h[0]; out->tris = h[1]; out->instances = h[2]; out->rays = h[3];
  return 0;
}

int rtc_generate_primary(rtc_context* ctx, const rt_SystemData* sys, uint32_t launchWidth, uint32_t launchHeight,
                         int iteration, uint64_t rays)
{
  if (!sys) RTC_FAIL("sys is null");
  RTC_CUDA(cudaSetDevice(ctx->device));
  return launch_generate_primary(ctx, *sys, launchWidth, launchHeight, iteration, (rtc_ray*)(uintptr_t)rays);
}

int rtc_composite(rtc_context* ctx, const rt_CompositorData* args)
{
  if (!args) RTC_FAIL("args is null");
  RTC_CUDA(cudaSetDevice(ctx->device));
  return launch_composite(ctx, *args);
}

int rtc_tonemap(rtc_context* ctx, const rt_TonemapperParams* params, uint64_t rgba, uint64_t rgb8, uint64_t numPixels)
{
  if (!params) RTC_FAIL("params is null");
  RTC_CUDA(cudaSetDevice(ctx->device));
  return launch_tonemap(ctx, *params, (const float4*)(uintptr_t)rgba, (uint8_t*)(uintptr_t)rgb8, numPixels);
}

int rtc_probe_math(rtc_context* ctx, int fn, const float* x, const float* y, float* out, uint32_t n)
{
  if (!ctx || !x || !y || !out) RTC_FAIL("null argument");
  if (fn < 0 || fn >= RTC_MATH_COUNT) RTC_FAIL("unknown function");
  if (n == 0) return 0;
  RTC_CUDA(cudaSetDevice(ctx->device));
  float* d = nullptr;
  RTC_CUDA(cudaMalloc(&d, (size_t)n * 3 * sizeof(float)));
  const bool ok = cudaMemcpyAsync(d, x, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
                  cudaMemcpyAsync(d + n, y, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
                  launch_probe_math(ctx, fn, d, d + n, d + 2 * (size_t)n, n) == 0 &&
                  cudaMemcpyAsync(out, d + 2 * (size_t)n, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
                  cudaStreamSynchronize(ctx->stream) == cudaSuccess;
  cudaFree(d);
  if (!ok) RTC_FAIL("device error");
  return 0;
}

int rtc_stats_get(rtc_context* ctx, rtc_stats* out)
{
  if (!out) RTC_FAIL("out is null");
  RTC_CUDA(cudaSetDevice(ctx->device));
  uint64_t h[4];
  RTC_CUDA(cudaMemcpyAsync(h, ctx->d_stats, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  RTC_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->traceTimed)
  {
    float ms = 0.0f;
    RTC_CUDA(cudaEventElapsedTime(&ms, ctx->evA, ctx->evB));
    ctx->lastTraceMs = ms;
  }
  out->radianceRays = h[0]; out->shadowRays = h[1]; out->pathSamples = h[2];
  out->kernelLaunches = ctx->kernelLaunches; out->lastTraceMs = ctx->lastTraceMs;
  return read_stack_overflows(ctx, &out->stackOverflows);
}

int rtc_stats_reset(rtc_context* ctx)
{
  RTC_CUDA(cudaSetDevice(ctx->device));
  RTC_CUDA(cudaMemsetAsync(ctx->d_stats, 0, 4 * sizeof(uint64_t), ctx->stream));
  ctx->kernelLaunches = 0;
  return 0;
}

} // extern "C"
