/*
 * ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.  Runs the reference's host-compiled device programs (see
 * oracle/ref_shim/optix.h) one launch index at a time: the job optixLaunch + the OptiX runtime do on a GPU
 * (apps/rtigo3/src/DeviceSingleGPU.cpp:164; program groups and the direct-callable table, src/Device.cpp:634-790).
 * Traversal is delegated to the scalar oracle built with libm transcendentals (liborc_libm.so).
 */
#include <cstdio>
#include <cstring>
#include <vector>

#include <optix.h>

#include "system_data.h"     // the reference's own headers, found through -I /root/reference/apps/rtigo3/shaders
#include "per_ray_data.h"
#include "light_definition.h"
#include "material_definition.h"
#include "shader_common.h"

#include "rt_oracle.h"

extern "C" { SystemData sysData; }

// programs defined by the reference's translation units
extern "C" void __raygen__path_tracer();
extern "C" void __raygen__path_tracer_local_copy();
extern "C" void __closesthit__radiance();
extern "C" void __anyhit__shadow();
extern "C" void __anyhit__radiance_cutout();
extern "C" void __anyhit__shadow_cutout();
extern "C" void __miss__env_null();
extern "C" void __miss__env_constant();
extern "C" void __miss__env_sphere();
extern "C" void __direct_callable__pinhole(const float2, const float2, const float2, float3&, float3&);
extern "C" void __direct_callable__fisheye(const float2, const float2, const float2, float3&, float3&);
extern "C" void __direct_callable__sphere(const float2, const float2, const float2, float3&, float3&);
extern "C" void __direct_callable__light_env_constant(float3 const&, const float2, LightSample&);
extern "C" void __direct_callable__light_env_sphere(float3 const&, const float2, LightSample&);
extern "C" void __direct_callable__light_parallelogram(float3 const&, const float2, LightSample&);
extern "C" void __direct_callable__sample_brdf_diffuse(MaterialDefinition const&, State const&, PerRayData*);
extern "C" float4 __direct_callable__eval_brdf_diffuse(MaterialDefinition const&, State const&, PerRayData* const, float3 const&);
extern "C" void __direct_callable__sample_brdf_specular(MaterialDefinition const&, State const&, PerRayData*);
extern "C" float4 __direct_callable__eval_brdf_specular(MaterialDefinition const&, State const&, PerRayData* const, float3 const&);
extern "C" void __direct_callable__sample_bsdf_specular(MaterialDefinition const&, State const&, PerRayData*);
extern "C" void __direct_callable__sample_brdf_ggx_smith(MaterialDefinition const&, State const&, PerRayData*);
extern "C" float4 __direct_callable__eval_brdf_ggx_smith(MaterialDefinition const&, State const&, PerRayData* const, float3 const&);
extern "C" void __direct_callable__sample_bsdf_ggx_smith(MaterialDefinition const&, State const&, PerRayData*);

namespace {

struct EnvTexture { const float* texels; unsigned int width, height; };

struct Runtime
{
  const orc_scene* scene = nullptr;
  int miss = 0;
  std::vector<GeometryInstanceData> sbt;      // one record per instance
  std::vector<float4> objectToWorld;          // 3 per instance
  std::vector<float4> worldToObject;          // 3 per instance
  void* callables[15] = {};
  EnvTexture env = { nullptr, 0, 0 };
  // launch state
  uint3 launchIndex, launchDim;
  // current trace state
  unsigned int payload0 = 0, payload1 = 0;
  unsigned int hitInstance = 0, hitPrimitive = 0;
  float2 barycentrics;
  float rayTmax = 0.0f;
  bool ignored = false;       // set by optixIgnoreIntersection inside an any-hit program
  bool hasCutout = false;     // some material carries a cutout texture (decided per ref_render_rows call)
};

Runtime g;

} // namespace

uint3 optixGetLaunchIndex() { return g.launchIndex; }
uint3 optixGetLaunchDimensions() { return g.launchDim; }
unsigned int optixGetPayload_0() { return g.payload0; }
unsigned int optixGetPayload_1() { return g.payload1; }
CUdeviceptr optixGetSbtDataPointer() { return (CUdeviceptr)(uintptr_t)&g.sbt[g.hitInstance]; }
unsigned int optixGetPrimitiveIndex() { return g.hitPrimitive; }
float2 optixGetTriangleBarycentrics() { return g.barycentrics; }
float optixGetRayTmax() { return g.rayTmax; }
OptixTraversableHandle optixGetTransformListHandle(unsigned int) { return (OptixTraversableHandle)g.hitInstance; }
const float4* optixGetInstanceTransformFromHandle(OptixTraversableHandle h) { return &g.objectToWorld[3 * (size_t)h]; }
const float4* optixGetInstanceInverseTransformFromHandle(OptixTraversableHandle h) { return &g.worldToObject[3 * (size_t)h]; }
void optixTerminateRay() {}
void optixIgnoreIntersection() { g.ignored = true; }
unsigned int optixGetExceptionCode() { return 0; }
void* ref_callable(unsigned int sbtIndex) { return sbtIndex < 15 ? g.callables[sbtIndex] : nullptr; }

// The software texture fetch this repository DEFINES for the environment map (bilinear, wrap u, clamp v), identical to
// env_lookup in oracle/rt_oracle.c and csrc/shade.cuh; the reference used the texture unit (miss.cu:90, light_sample.cu:147).
// Material textures (closesthit.cu:235, anyhit.cu:70, :119): the handle is the address of {uint32 w, h, 0, 0} + RGBA32F
// texels and the fetch is the one oracle/rt_oracle.c DEFINES (bilinear, wrap/wrap).
template <> float4 tex2D<float4>(cudaTextureObject_t texture, float u, float v)
{
  if ((uintptr_t)texture != (uintptr_t)&g.env)
  {
    float rgb[3];
    orc_tex2d((uint64_t)texture, u, v, rgb);
    return make_float4(rgb[0], rgb[1], rgb[2], 1.0f);
  }
  const EnvTexture* t = reinterpret_cast<const EnvTexture*>((uintptr_t)texture);
  const int W = (int)t->width, H = (int)t->height;
  const float x = u * (float)W - 0.5f, y = v * (float)H - 0.5f;
  const float fx = floorf(x), fy = floorf(y);
  const float ax = x - fx, ay = y - fy;
  int x0 = (int)fx % W; if (x0 < 0) x0 += W;
  int x1 = x0 + 1; if (x1 >= W) x1 = 0;
  int y0 = (int)fy, y1 = y0 + 1;
  if (y0 < 0) y0 = 0; if (y0 > H - 1) y0 = H - 1;
  if (y1 < 0) y1 = 0; if (y1 > H - 1) y1 = H - 1;
  const float* t00 = &t->texels[4 * ((size_t)y0 * W + x0)];
  const float* t10 = &t->texels[4 * ((size_t)y0 * W + x1)];
  const float* t01 = &t->texels[4 * ((size_t)y1 * W + x0)];
  const float* t11 = &t->texels[4 * ((size_t)y1 * W + x1)];
  float c[4];
  for (int k = 0; k < 4; ++k)
  {
    const float a = t00[k] + ax * (t10[k] - t00[k]);
    const float b = t01[k] + ax * (t11[k] - t01[k]);
    c[k] = a + ay * (b - a);
  }
  return make_float4(c[0], c[1], c[2], c[3]);
}

// Next candidate intersection after `skip` in the canonical order (t, instance, primitive); false when there is none.
static bool next_candidate(const orc_ray& ray, const orc_hit* skip, orc_hit& hit)
{
  if (skip) orc_trace_closest_after(g.scene, &ray, skip->t, skip->inst, skip->prim, &hit);
  else      orc_trace_closest(g.scene, &ray, 1, 0, &hit, nullptr);
  if (hit.inst == 0xffffffffu) return false;
  g.hitInstance = hit.inst; g.hitPrimitive = hit.prim; g.barycentrics = make_float2(hit.u, hit.v); g.rayTmax = hit.t;
  return true;
}

// The instance's hit records are the cutout ones when its material has a cutout texture (src/Device.cpp:1503-1513).
static bool cutout_records(unsigned int instance)
{
  return sysData.materialDefinitions[g.sbt[instance].materialIndex].textureCutout != 0;
}

// optixTrace: closest hit -> __closesthit__radiance or the miss program; DISABLE_CLOSESTHIT (shadow rays) -> any hit
// runs __anyhit__shadow once, a miss does nothing (the shadow miss program is null, src/Device.cpp:674-678).
// With cutout materials in the scene the reference's any-hit programs run per candidate, in the canonical order this
// repository defines (closest candidate first; see "Cutout opacity" in rt_oracle.c).
void ref_trace(OptixTraversableHandle, float3 origin, float3 direction, float tmin, float tmax, float,
               unsigned int, unsigned int rayFlags, unsigned int, unsigned int, unsigned int, unsigned int& p0, unsigned int& p1)
{
  orc_ray ray = { origin.x, origin.y, origin.z, tmin, direction.x, direction.y, direction.z, tmax };
  const unsigned int save0 = g.payload0, save1 = g.payload1;
  g.payload0 = p0; g.payload1 = p1;
  if (g.hasCutout)
  {
    orc_hit hit, skip; const orc_hit* after = nullptr;
    const bool shadowRay = (rayFlags & OPTIX_RAY_FLAG_DISABLE_CLOSESTHIT) != 0;
    bool accepted = false;
    while (next_candidate(ray, after, hit))
    {
      g.ignored = false;
      if (shadowRay) { if (cutout_records(hit.inst)) __anyhit__shadow_cutout(); else __anyhit__shadow(); }
      else if (cutout_records(hit.inst)) __anyhit__radiance_cutout();
      if (!g.ignored) { accepted = true; break; }
      skip = hit; after = &skip;
    }
    if (!shadowRay)
    {
      if (accepted) __closesthit__radiance();          // g.hit* still describe the accepted candidate
      else if (g.miss == 2) __miss__env_sphere();
      else if (g.miss == 1) __miss__env_constant();
      else                  __miss__env_null();
    }
  }
  else if (rayFlags & OPTIX_RAY_FLAG_DISABLE_CLOSESTHIT)
  {
    uint8_t occluded = 0;
    orc_trace_any(g.scene, &ray, 1, 0, &occluded, nullptr);
    if (occluded) __anyhit__shadow();
  }
  else
  {
    orc_hit hit;
    orc_trace_closest(g.scene, &ray, 1, 0, &hit, nullptr);
    if (hit.inst != 0xffffffffu)
    {
      g.hitInstance = hit.inst; g.hitPrimitive = hit.prim; g.barycentrics = make_float2(hit.u, hit.v); g.rayTmax = hit.t;
      __closesthit__radiance();
    }
    else if (g.miss == 2) __miss__env_sphere();
    else if (g.miss == 1) __miss__env_constant();
    else                  __miss__env_null();
  }
  p0 = g.payload0; p1 = g.payload1;
  g.payload0 = save0; g.payload1 = save1;
}

extern "C" {

// instances: per instance {attributes ptr, indices ptr, material, light}; transforms: 12 floats each.
// env: may be null.  All arrays must outlive ref_render calls.
void ref_setup(const orc_scene* scene, int miss, int numInstances, const void* const* attributes, const void* const* indices,
               const int* materialIndex, const int* lightIndex, const float* transforms,
               const void* cameras, const void* lights, const void* materials,
               const float* envTexels, unsigned int envWidth, unsigned int envHeight, const float* envCdfU, const float* envCdfV, float envIntegral)
{
  g.scene = scene; g.miss = miss;
  g.sbt.resize((size_t)numInstances); g.objectToWorld.resize(3 * (size_t)numInstances); g.worldToObject.resize(3 * (size_t)numInstances);
  for (int i = 0; i < numInstances; ++i)
  {
    g.sbt[i].attributes = (CUdeviceptr)(uintptr_t)attributes[i];
    g.sbt[i].indices = (CUdeviceptr)(uintptr_t)indices[i];
    g.sbt[i].materialIndex = materialIndex[i];
    g.sbt[i].lightIndex = lightIndex[i];
    float inv[12];
    orc_scene_get_inverse(scene, i, inv);
    for (int r = 0; r < 3; ++r)
    {
      const float* m = &transforms[12 * (size_t)i + 4 * r];
      g.objectToWorld[3 * (size_t)i + r] = make_float4(m[0], m[1], m[2], m[3]);
      g.worldToObject[3 * (size_t)i + r] = make_float4(inv[4 * r], inv[4 * r + 1], inv[4 * r + 2], inv[4 * r + 3]);
    }
  }
  // the direct-callable table in SBT order (src/Device.cpp:690-790; function_indices.h)
  g.callables[0] = (void*)&__direct_callable__pinhole;
  g.callables[1] = (void*)&__direct_callable__fisheye;
  g.callables[2] = (void*)&__direct_callable__sphere;
  g.callables[3] = (miss == 2) ? (void*)&__direct_callable__light_env_sphere : (void*)&__direct_callable__light_env_constant;
  g.callables[4] = (void*)&__direct_callable__light_parallelogram;
  g.callables[5] = (void*)&__direct_callable__sample_brdf_diffuse;
  g.callables[6] = (void*)&__direct_callable__eval_brdf_diffuse;
  g.callables[7] = (void*)&__direct_callable__sample_brdf_specular;
  g.callables[8] = (void*)&__direct_callable__eval_brdf_specular;
  g.callables[9] = (void*)&__direct_callable__sample_bsdf_specular;
  g.callables[10] = (void*)&__direct_callable__eval_brdf_specular;
  g.callables[11] = (void*)&__direct_callable__sample_brdf_ggx_smith;
  g.callables[12] = (void*)&__direct_callable__eval_brdf_ggx_smith;
  g.callables[13] = (void*)&__direct_callable__sample_bsdf_ggx_smith;
  g.callables[14] = (void*)&__direct_callable__eval_brdf_specular;

  std::memset(&sysData, 0, sizeof(sysData));
  sysData.cameraDefinitions = (CameraDefinition*)cameras;
  sysData.lightDefinitions = (LightDefinition*)lights;
  sysData.materialDefinitions = (MaterialDefinition*)materials;
  g.env.texels = envTexels; g.env.width = envWidth; g.env.height = envHeight;
  sysData.envTexture = (cudaTextureObject_t)(uintptr_t)&g.env;
  sysData.envCDF_U = const_cast<float*>(envCdfU);
  sysData.envCDF_V = const_cast<float*>(envCdfV);
  sysData.envWidth = envWidth; sysData.envHeight = envHeight; sysData.envIntegral = envIntegral;
}

// values: the non-pointer fields of SystemData as this repository's rt_SystemData carries them (same layout).
// rowStep / rowOffset restrict the work to launch rows y with y % rowStep == rowOffset, so that several PROCESSES (the
// reference keeps its launch parameters in one global, so threads cannot share it) can split a frame between them.
void ref_render_rows(const rt_SystemData* values, unsigned int launchWidth, unsigned int launchHeight, int localCopy,
                     int iterFirst, int iterCount, int rowStep, int rowOffset, float* buffer);

void ref_render(const rt_SystemData* values, unsigned int launchWidth, unsigned int launchHeight, int localCopy,
                int iterFirst, int iterCount, float* buffer)
{
  ref_render_rows(values, launchWidth, launchHeight, localCopy, iterFirst, iterCount, 1, 0, buffer);
}

void ref_render_rows(const rt_SystemData* values, unsigned int launchWidth, unsigned int launchHeight, int localCopy,
                     int iterFirst, int iterCount, int rowStep, int rowOffset, float* buffer)
{
  if (rowStep < 1) rowStep = 1;
  sysData.resolution = make_int2(values->resolution.x, values->resolution.y);
  sysData.tileSize = make_int2(values->tileSize.x, values->tileSize.y);
  sysData.tileShift = make_int2(values->tileShift.x, values->tileShift.y);
  sysData.pathLengths = make_int2(values->pathLengths.x, values->pathLengths.y);
  sysData.deviceCount = values->deviceCount; sysData.deviceIndex = values->deviceIndex; sysData.distribution = values->distribution;
  sysData.samplesSqrt = values->samplesSqrt; sysData.sceneEpsilon = values->sceneEpsilon; sysData.clockScale = values->clockScale;
  sysData.lensShader = values->lensShader; sysData.numCameras = values->numCameras; sysData.numMaterials = values->numMaterials;
  sysData.numLights = values->numLights; sysData.envRotation = values->envRotation;
  sysData.outputBuffer = (CUdeviceptr)(uintptr_t)buffer;
  sysData.texelBuffer = (CUdeviceptr)(uintptr_t)buffer;
  g.launchDim = make_uint3(launchWidth, launchHeight, 1u);
  g.hasCutout = false;
  for (int m = 0; m < sysData.numMaterials; ++m) if (sysData.materialDefinitions[m].textureCutout != 0) g.hasCutout = true;
  for (int it = iterFirst; it < iterFirst + iterCount; ++it)
  {
    sysData.iterationIndex = it;
    for (unsigned int y = 0; y < launchHeight; ++y)
    {
      if ((int)(y % (unsigned int)rowStep) != rowOffset) continue;
      for (unsigned int x = 0; x < launchWidth; ++x)
      {
        g.launchIndex = make_uint3(x, y, 0u);
        if (localCopy) __raygen__path_tracer_local_copy(); else __raygen__path_tracer();
      }
    }
  }
}

int ref_sizeof_system_data(void) { return (int)sizeof(SystemData); }
int ref_sizeof_per_ray_data(void) { return (int)sizeof(PerRayData); }

// known-answer access to the reference's RNG (random_number_generators.h)
}
#include "random_number_generators.h"
extern "C" {
unsigned int ref_tea4(unsigned int a, unsigned int b) { return tea<4>(a, b); }
float ref_rng(unsigned int* state) { return rng(*state); }
}
