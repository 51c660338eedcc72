/*
 * rt_oracle.h -- TEST INFRASTRUCTURE ONLY.  Scalar CPU restatement of the rtigo3 hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (tweeker_raytracer_b200/) never links or calls it.
 *
 * What it restates (reference file:line, relative to /root/reference/apps/rtigo3/):
 *   raygeneration      shaders/raygeneration.cu:42-149 (integrator), :152-164 (distribute),
 *                      :167-256 (path_tracer), :259-344 (path_tracer_local_copy)
 *   closest hit        shaders/closesthit.cu:126-305
 *   any hit            shaders/anyhit.cu:84-91 (shadow), :46-80 and :94-132 (cutout opacity, canonical candidate order)
 *   miss               shaders/miss.cu:41-109
 *   lens shaders       shaders/lens_shader.cu:40-99
 *   light sampling     shaders/light_sample.cu:42-177
 *   BSDFs              shaders/bxdf_diffuse.cu:39-95, bxdf_specular.cu:42-134, bxdf_ggx_smith.cu:45-319
 *   RNG                shaders/random_number_generators.h:39-78
 *   helpers            shaders/shader_common.h:47-187, vector_math.h:130-150,436-446,547-608
 *   compositor         shaders/compositor.cu:38-65
 *   tonemapper         src/Application.cpp:2262-2295
 *
 * What it DEFINES because the reference has no source for it (OptiX 7.0 / libnvoptix, pinned by
 * apps/CMake/FindOptiX7.cmake; optixTrace at raygeneration.cu:84 and closesthit.cu:281;
 * optixAccelBuild at src/Device.cpp:1401,1478): the ray/triangle test (watertight, after
 * Woop, Benthin, Wald 2013), the closest-hit tie rule, the instance inverse, and a BVH.
 * PARITY FOR THAT PART IS UNPINNED BY THE REFERENCE (it holds no tests or golden vectors,
 * SURVEY.md section 4); it is pinned here by brute force == BVH and analytic known answers.
 *
 * Pins for the restated part: oracle/_ref (the reference's own shader sources compiled for the
 * host behind a small shim, see oracle/Makefile) must agree with this oracle built with
 * RT_MATH_LIBM; tests/test_oracle_vs_reference.py checks that and tests/golden/ holds vectors
 * generated from it.
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H

#include "rtigo3_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;

/* One ray: origin, tmin, direction, tmax (8 floats), same packing as the core's rtc_trace(). */
typedef struct { float ox, oy, oz, tmin, dx, dy, dz, tmax; } orc_ray;

/* One closest hit.  t < 0 means miss (inst = prim = 0xffffffff). */
typedef struct { float t, u, v; uint32_t inst; uint32_t prim; } orc_hit;

typedef struct {
  uint64_t radianceRays;
  uint64_t shadowRays;
  uint64_t pathSamples;
  uint64_t nodesVisited;     /* BVH nodes popped (own binary BVH, or the wide BVH in orc_trace_wide) */
  uint64_t trisTested;
  uint64_t instancesEntered;
} orc_stats;

orc_scene* orc_scene_create(void);
void       orc_scene_destroy(orc_scene* s);

/* Geometry = one GAS input of the reference (Device::createGeometry, src/Device.cpp:1333-1425). Data is copied. */
int  orc_scene_add_geometry(orc_scene* s, const rt_TriangleAttributes* attrs, uint32_t numVerts,
                            const uint32_t* indices, uint32_t numTris);
/* Instance = Device::createInstance (src/Device.cpp:1427-1443): instanceId is the call order. */
int  orc_scene_add_instance(orc_scene* s, const float transform[12], int geometry, int material, int light);
void orc_scene_set_materials(orc_scene* s, const rt_MaterialDefinition* m, int n);
void orc_scene_set_lights(orc_scene* s, const rt_LightDefinition* l, int n);
void orc_scene_set_camera(orc_scene* s, const rt_CameraDefinition* c);
void orc_scene_set_env(orc_scene* s, const float* rgba, uint32_t w, uint32_t h,
                       const float* cdfU, const float* cdfV, float integral);
/* Builds the oracle's own BVHs and the instance inverses. */
void orc_scene_commit(orc_scene* s);
/* Copies out the 3x4 world->object matrix of an instance (after commit). */
void orc_scene_get_inverse(const orc_scene* s, int instance, float out[12]);

/* mode: 0 = oracle BVH, 1 = brute force over every triangle of every instance. */
void orc_trace_closest(const orc_scene* s, const orc_ray* rays, uint64_t n, int mode, orc_hit* hits, orc_stats* stats);
void orc_trace_any(const orc_scene* s, const orc_ray* rays, uint64_t n, int mode, uint8_t* occluded, orc_stats* stats);

/* Closest hit AFTER the key (skipT, skipInst, skipPrim) in the canonical candidate order (t, instance, primitive): the
 * step of the ordered any-hit processing (cutout opacity, anyhit.cu:46-132; see "Cutout opacity" in rt_oracle.c). */
void orc_trace_closest_after(const orc_scene* s, const orc_ray* ray, float skipT, uint32_t skipInst, uint32_t skipPrim, orc_hit* hit);
/* The material-texture fetch DEFINED by this repository (bilinear, wrap/wrap; handle = address of {uint32 w, h, 0, 0}
 * followed by w*h RGBA32F texels).  out = rgb. */
void orc_tex2d(uint64_t handle, float u, float v, float out[3]);

/* Primary rays as the raygeneration program makes them for iteration `iteration` (one per launch index,
 * row-major launchWidth x launchHeight); rays of skipped launch indices get tmax = -1. */
void orc_generate_primary(const orc_scene* s, const rt_SystemData* sys, uint32_t launchWidth, uint32_t launchHeight,
                          int iteration, orc_ray* rays);

/*
 * Render iterations [iterFirst, iterFirst+iterCount) into `buffer` (float4 per pixel) exactly as
 * __raygen__path_tracer (localCopy = 0: buffer is resolution-sized outputBuffer) or
 * __raygen__path_tracer_local_copy (localCopy = 1: buffer is launchWidth x launchHeight texelBuffer) would.
 * Only sys->{resolution,tileSize,tileShift,pathLengths,deviceCount,deviceIndex,distribution,sceneEpsilon,
 * lensShader,numLights,envRotation} are read; pointers in sys are ignored (scene data comes from `s`).
 * miss = RT_MISS_*.  rowStep/rowOffset restrict the work to launch rows y with y % rowStep == rowOffset
 * (bounded samples for the CPU baseline).  threads <= 0 means one per online core.
 */
void orc_render(const orc_scene* s, const rt_SystemData* sys, int miss, uint32_t launchWidth, uint32_t launchHeight,
                int localCopy, int iterFirst, int iterCount, int rowStep, int rowOffset, int threads,
                float* buffer, orc_stats* stats);

/* Radiance of single path samples (no accumulation): out[3*i..] for pixel list (x,y) pairs; used for diagnostics. */
void orc_path_radiance(const orc_scene* s, const rt_SystemData* sys, int miss, uint32_t launchWidth,
                       const uint32_t* launchXY, uint64_t n, int iteration, float* out, orc_stats* stats);

/* shaders/compositor.cu:38-65 for one source device. */
void orc_composite(const rt_CompositorData* args, const float* tileBuffer, float* outputBuffer);

/* src/Application.cpp:2262-2295. in: float4 per pixel, out: 3 bytes per pixel. */
void orc_tonemap(const rt_TonemapperParams* p, const float* rgba, uint8_t* rgb, uint64_t numPixels);

/* RNG known-answer access (random_number_generators.h:39-78). */
uint32_t orc_tea4(uint32_t v0, uint32_t v1);
float    orc_rng(uint32_t* state);

int orc_online_cores(void);
/* 1 when built with RT_MATH_LIBM (libm transcendentals), 0 for the pinned arithmetic. */
int orc_uses_libm(void);

#ifdef __cplusplus
}
#endif
#endif
