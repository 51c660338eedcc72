// rtc_internal.h -- structures shared by the host side of librtcore and its sm_100a kernels.
//
// HBM layout of one scene (everything the traversal touches is 16-byte addressable so every
// fetch is one LDG.128):
//   per GAS    nodes  uint4[5] per wide node (80 B): 8 children, quantised boxes    -> Node8 below (root = node 0)
//              tris   float4[3] per triangle (48 B): v0.xyz|prim id, v1.xyz|0, v2.xyz|0   (leaf order)
//   per scene  tlasNodes  the same 80 B nodes over instance bounds (root = node 0)
//              tlasLeaves uint32 per instance-level leaf slot -> instance id
//              instances  float4[4] per instance (64 B): world->object rows 0..2, {GAS nodes ptr, GAS tris ptr}
//   shading tables (used by shade only): objectToWorld float4[3], rt_GeometryInstanceData per instance
// Node and triangle indices are relative to their own GAS, so a GAS is built once (on the host or by the
// GPU builder, straight into device memory) and shared by any number of instances and scenes.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "rtc_core.h"

// 8-wide node with child boxes quantised to 8 bits per plane on a per-node power-of-two grid
// (after Ylitie, Karras, Laine, "Efficient Incoherent Ray Traversal on GPUs Through Compressed
// Wide BVHs", HPG 2017).  Child i occupies slot i; slots are assigned by octant so that
// slot ^ (7 ^ rayOctant) is a front-to-back priority.
//   empty slot : imask bit clear, meta == 0
//   inner child: imask bit set; its node is childBase + popc(imask & ((1 << slot) - 1))
//   leaf child : imask bit clear, meta = (count << 5) | offset; primitives triBase + offset .. + count - 1
struct Node8
{
  float    px, py, pz;
  uint8_t  ex, ey, ez;      // biased exponents: grid step = 2^(e - 127)
  uint8_t  imask;
  uint32_t childBase;
  uint32_t triBase;
  uint8_t  meta[8];
  uint8_t  qlox[8], qloy[8];
  uint8_t  qloz[8], qhix[8];
  uint8_t  qhiy[8], qhiz[8];
};
static_assert(sizeof(Node8) == 80, "wide node is 80 bytes");

// Device-visible scene descriptor; SystemData::topObject points at one of these (in device memory).
struct SceneDesc
{
  const uint4*    tlasNodes;
  const uint32_t* tlasLeaves;
  const float4*   instances;
  const float4*   objectToWorld;
  const rt_GeometryInstanceData* geomInst;
  const uint32_t* instFlags;      // RTC_INSTANCE_* per instance (which hit records the instance uses)
  uint32_t        numInstances;
  uint32_t        numTlasNodes;
  uint32_t        numTlasLeaves;
  uint32_t        pad0;
};

// Host-side result of a BVH build over generic primitive boxes.
struct WideBvh
{
  std::vector<Node8>    nodes;      // nodes[0] is the root
  std::vector<uint32_t> primOrder;  // leaf order -> input primitive index
  float lo[3], hi[3];               // bounds of everything
};

struct PrimBox { float lo[3], hi[3]; };

// bvh_build_host.cpp: binned-SAH binary build, greedy collapse to 8-wide, octant slot assignment, quantisation.
// leafMax: primitives per leaf child (1..3).  Triangles: 2 (rtc_gas_build).  Instances: 1 -- entering an instance costs about three node
// visits (transform, shear constants, the GAS root), so a leaf box shared by two or three instances is never worth it
// (geometry scene: 1.30 -> 1.06 instance entries per ray, +7.5 % samples/s).
void build_wide_bvh_host(const PrimBox* prims, uint32_t numPrims, WideBvh& out, uint32_t leafMax = 3, bool instanceLevel = false);

// accel_host.cpp: the host halves of rtc_gas_build (host SAH builder) and rtc_ias_build -- no CUDA call in any of them -- shared
// with the host-only twin of the two builds (rtc_host_gas_build / rtc_host_ias_build, include/rtc_core.h).
void invert_3x4(const float m[12], float out[12]);
bool gas_assemble_host(const uint8_t* verts, uint32_t strideBytes, uint32_t numVerts, const uint32_t* idx, uint32_t numTris,
                       WideBvh& bvh, std::vector<float4>& tris);
void instance_bounds_host(const float transform[12], const uint8_t* verts, uint32_t strideBytes, uint32_t numVerts, PrimBox& out);
void instance_box_finish(const float transform[12], const float gasLo[3], const float gasHi[3], bool tight, bool emptyGas, PrimBox& b);
bool instance_bounds_tight();      // false with RTC_INSTANCE_BOUNDS=box
void tlas_build_host(const PrimBox* boxes, uint32_t numInstances, WideBvh& bvh);

struct GasRecord
{
  void*    d_nodes = nullptr;              // Node8[numNodes]
  void*    d_tris = nullptr;               // float4[3 * numTris]
  uint32_t numNodes = 0, numTris = 0;
  float    lo[3] = { 0, 0, 0 }, hi[3] = { 0, 0, 0 };
  uint64_t attributes = 0, indices = 0;    // caller-owned inputs (kept for the per-instance shading table)
  uint32_t strideBytes = 0, numVerts = 0;
  double   buildMs = 0.0;
  int      builder = 0;                    // RTC_BUILD_HOST_SAH or RTC_BUILD_GPU_LBVH
};

struct SceneRecord
{
  SceneDesc desc{};          // host copy
  SceneDesc* d_desc = nullptr;
  void *d_tlasNodes = nullptr, *d_instances = nullptr, *d_tlasLeaves = nullptr, *d_o2w = nullptr, *d_geomInst = nullptr, *d_instFlags = nullptr;
  std::vector<uint32_t> instFlags;         // host copy of the per-instance flags
  uint32_t numCutout = 0;                  // instances using the cutout hit records: > 0 selects the ordered any-hit path
  bool     albedoTextures = false;         // MaterialDefinition.textureAlbedo may be non-zero: selects the textured shade kernels
  std::vector<float> inverses;             // 12 floats per instance (host copy of the world->object matrices)
  std::vector<uint32_t> instGas;           // GAS handle per instance
  uint64_t totalNodes = 0, totalTris = 0;  // unique GAS nodes / triangles referenced + instance level
  uint32_t numGas = 0;
  double   gasBuildMs = 0.0, iasBuildMs = 0.0;
};

// Wavefront state of one launch batch (device pointers, SoA, sized for `capacity` paths).
struct WavefrontBuffers
{
  uint64_t capacity = 0;
  float4 *rayOrg = nullptr, *rayDir = nullptr;     // per path: next radiance ray (org.xyz,tmin) (dir.xyz,tmax)
  float4 *hit = nullptr;                           // per path: t,u,v,prim bits
  uint32_t *hitInst = nullptr;                     // per path
  float4 *throughput = nullptr;                    // per path: T.xyz, pdf
  float4 *radiance = nullptr;                      // per path: L.xyz, flags bits
  uint4  *misc = nullptr;                          // per path: seed, depth, stackIdx, pixel index
  float4 *absStack = nullptr;                      // per path x 4: nested-volume stack
  float4 *shadowOrg = nullptr, *shadowDir = nullptr, *shadowContrib = nullptr;  // per path
  uint32_t *queueA = nullptr, *queueB = nullptr, *shadowQueue = nullptr;
  uint32_t *bins = nullptr; uint32_t binStride = 0;  // per shade class: path ids binned after extend (class c at bins + c * binStride)
  uint32_t *counters = nullptr;                    // [0..63] extend counts per depth, [64..127] shadow counts per depth,
                                                   // [128..191] extend ray cursors, [192..255] connect ray cursors
  void* base = nullptr;
  // ordered any-hit processing (scenes with cutout materials only): two ping-pong queues of paths whose closest candidate
  // was ignored, and their counters {count A, count B, cursor}; allocated on first use
  uint32_t *cutQueue[2] = { nullptr, nullptr };
  uint32_t *cutCounters = nullptr;
  uint64_t cutCapacity = 0;
  void* cutBase = nullptr;
};

// RTC_SCHEDULE_GROUP / RTC_SCHEDULE_ONE_TRI / RTC_SCHEDULE_TWO_TRI (rtc_core.h): how the lane-owned traversal driver times its triangle tests
// (trace.cuh Traversal::step); hits and work counters do not depend on it.
// Run-time choice between the schedules.  The capped ones were written when no GPU was left to measure them on, so the library
// measures them itself: of the first batches of a context that are large enough to time (>= 1 Mi paths), one is a warm-up and
// the next four run group / one triangle / two triangles / group, each between two events; a capped schedule serves every later
// launch only if it beats the FASTER of the two group batches by 3 % (the faster capped one if both do).  Results are
// bit-identical under all of them, so the choice never shows in a frame.  RTC_TRACE_SCHEDULE=group|onetri|twotri fixes it (auto
// is the default); rtc_trace_schedule_get reports what happened.
struct ScheduleTuner
{
  enum State { WARMUP = 0, TIMING = 1, PENDING = 2, DONE = 3 };
  static constexpr int kSlots = 4;  // timed batches: group, one triangle, two triangles, group again
  int         state = WARMUP;
  int         slot = 0;             // TIMING: the next timed batch
  cudaEvent_t ev[2 * kSlots] = {};  // begin / end of each timed batch
  uint64_t    paths = 0;            // size of the timed batches (all must be the same)
  int         restarts = 0;         // a batch of another size restarts the measurement; bounded
  float       ms[kSlots] = { 0.0f, 0.0f, 0.0f, 0.0f };
  bool        fixedByEnv = false;
};

struct rtc_context
{
  int device = 0;
  cudaStream_t stream = nullptr;
  int numSMs = 0;
  std::vector<GasRecord> gas;
  std::vector<SceneRecord*> scenes;
  WavefrontBuffers wf;
  uint64_t* d_stats = nullptr;        // device counters: radiance rays, shadow rays, path samples
  uint64_t kernelLaunches = 0;
  double lastTraceMs = 0.0;
  bool traceTimed = false;
  cudaEvent_t evA = nullptr, evB = nullptr;
  cudaEvent_t evTimerA = nullptr, evTimerB = nullptr;
  // per-kernel-class profile (rtc_profile_enable): pending event pairs are resolved at rtc_profile_get
  bool profiling = false;
  struct ProfileSpan { int cls; cudaEvent_t a, b; };
  std::vector<ProfileSpan> spans;
  std::vector<cudaEvent_t> eventPool;
  rtc_profile profile{};
  std::vector<void*> textures;                    // rtc_texture_create allocations still alive
  // side streams of the shade stage: the class-specialised kernels of one depth are independent of each other and mostly
  // small, so they run concurrently (fork from `stream` after k_bin, join before connect); created on first use
  cudaStream_t shadeStreams[6] = {};
  cudaEvent_t  shadeFork = nullptr, shadeJoin[6] = {};
  unsigned long long* d_launchCounts = nullptr;   // 3 x kTraceCountWords: extend, connect, rtc_trace_count
  uint32_t* d_cursor = nullptr;                   // ray cursor of the query kernels (rtc_trace_*)
  bool   primaryPackets = false;                  // primary rays by packet traversal (trace_packet.cuh): measured slower, RTC_PRIMARY_PACKETS=1 turns it on
  int    traceDriver = 0;                         // RTC_DRIVER_LANE or RTC_DRIVER_POOL: which traversal driver the launches use
  int    traceSchedule = 0;                       // RTC_SCHEDULE_*: how the lane-owned driver times its triangle tests (what the NEXT launch uses)
  ScheduleTuner tuner;                            // picks traceSchedule by timing one batch with each (kernels_shade.cu, "schedule tuner")
  void*  cutoutGraph = nullptr;                   // CutoutGraph (kernels_shade.cu): the device-side loop of the ordered any-hit rounds
  void*  d_poolScratch = nullptr;                 // global part of the ray pool's traversal stacks (trace_pool.cuh), grown on demand
  size_t poolScratchBytes = 0;
};

enum { RTC_DRIVER_LANE = 0, RTC_DRIVER_POOL = 1 };     // trace.cuh trace_stream / trace_pool.cuh trace_pool

// work counters of one traversal kernel: {nodes, tris, insts, rays}, then the ray pool's passes and occupied lanes per phase (N, T, I, F)
constexpr int kTraceCountWords = 12;

// RAII-less helpers used by the launchers: bracket one kernel launch with an event pair when profiling is on
int profile_begin(rtc_context* ctx, int cls);
int profile_end(rtc_context* ctx);

// RTC_BUILD_DEFAULT picks the GPU LBVH builder above this many triangles (host SAH quality below it)
constexpr uint32_t kGpuBuildThreshold = 1u << 20;
// bvh_build_gpu.cu: Morton LBVH on the device, straight into rec.d_nodes / rec.d_tris; fills numNodes, lo, hi
int build_gas_gpu(rtc_context* ctx, GasRecord& rec);

// bvh_build_gpu.cu: exact world bounds of every instance's transformed vertices (one CTA per instance)
int instance_bounds_gpu(rtc_context* ctx, const rtc_instance_desc* instances, uint32_t numInstances, PrimBox* boxes);

// error plumbing (rtc_api.cpp)
int rtc_set_error(const char* file, int line, const char* call, int code, const char* text);
#define RTC_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return rtc_set_error(__FILE__, __LINE__, #call, (int)e_, cudaGetErrorString(e_)); } while (0)
#define RTC_FAIL(text) return rtc_set_error(__FILE__, __LINE__, __func__, -1, text)

// kernel launchers (kernels_trace.cu / kernels_shade.cu); all enqueue on ctx->stream
int launch_trace_closest(rtc_context* ctx, const SceneDesc* d_scene, const rtc_ray* rays, uint64_t n, rtc_hit* hits);
int launch_trace_any(rtc_context* ctx, const SceneDesc* d_scene, const rtc_ray* rays, uint64_t n, uint32_t* occluded);
// counting variant: adds {nodes popped, triangles tested, instances entered, rays} to d_counts[0..3]
int launch_trace_count(rtc_context* ctx, const SceneDesc* d_scene, const rtc_ray* rays, uint64_t n, int anyHit, unsigned long long* d_counts);
// cursor: zero-initialised device counter the persistent warps hand rays out from (one per launch)
int launch_extend(rtc_context* ctx, const SceneDesc* d_scene, const WavefrontBuffers& wf, const uint32_t* queue, const uint32_t* count,
                  uint32_t* cursor, bool countWork);
int launch_connect(rtc_context* ctx, const SceneDesc* d_scene, const WavefrontBuffers& wf, const uint32_t* count, uint32_t* cursor, bool countWork);
// ordered any-hit processing (cutout materials): re-trace past an ignored candidate / shadow rays as closest-hit queries
int launch_extend_after(rtc_context* ctx, const SceneDesc* d_scene, const WavefrontBuffers& wf, const uint32_t* queue, const uint32_t* count, uint32_t* cursor);
int launch_connect_closest(rtc_context* ctx, const SceneDesc* d_scene, const WavefrontBuffers& wf, const uint32_t* queue, const uint32_t* count,
                           uint32_t* cursor, bool after);
int launch_generate_primary(rtc_context* ctx, const rt_SystemData& sys, uint32_t w, uint32_t h, int iteration, rtc_ray* rays);
int launch_wavefront(rtc_context* ctx, const rt_SystemData& sys, uint32_t w, uint32_t h, int raygen, int miss, int iterFirst, int iterCount,
                     int accumFirst, bool countWork);
int launch_composite(rtc_context* ctx, const rt_CompositorData& args);
int launch_tonemap(rtc_context* ctx, const rt_TonemapperParams& p, const float4* rgba, uint8_t* rgb, uint64_t n);
int launch_probe_math(rtc_context* ctx, int fn, const float* x, const float* y, float* out, uint32_t n);   // device pointers
int ensure_wavefront(rtc_context* ctx, uint64_t capacity, bool* outOfMemory = nullptr);
int ensure_pool_scratch(rtc_context* ctx, size_t warps, uint2** out);
void release_cutout_graph(rtc_context* ctx);
int read_stack_overflows(rtc_context* ctx, uint64_t* out);
// schedule tuner (kernels_shade.cu); tuner_finish blocks for the last timed batch when a decision is pending
void tuner_finish(rtc_context* ctx);
void tuner_release(rtc_context* ctx);
void preload_capped_trace_kernels();       // kernels_trace.cu: loads the capped-schedule kernels before they are timed (lazy module loading)
int read_stack_overflows_primary(rtc_context* ctx, uint64_t* out);   // the counter of the primary-ray extend kernel (kernels_shade.cu)
