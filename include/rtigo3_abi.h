/*
 * rtigo3_abi.h -- the data contract of the rtigo3 hot path, restated for the B200-native core.
 *
 * Everything in this header is a LAYOUT or a CONSTANT that the reference's host classes
 * (Application / Raytracer / Device) hand to the device programs.  The B200 core keeps them
 * byte-for-byte so it is a drop-in behind those classes:
 *
 *   SystemData            apps/rtigo3/shaders/system_data.h:40-90      192 B, align 16
 *   GeometryInstanceData  apps/rtigo3/shaders/system_data.h:94-100      24 B
 *   MaterialDefinition    apps/rtigo3/shaders/material_definition.h:37-56   64 B
 *   LightDefinition       apps/rtigo3/shaders/light_definition.h:43-61      80 B
 *   CameraDefinition      apps/rtigo3/shaders/camera_definition.h:34-40     48 B
 *   TriangleAttributes    apps/rtigo3/shaders/vertex_attributes.h:34-40     48 B
 *   CompositorData        apps/rtigo3/shaders/compositor_data.h:34-48       56 B
 *   flags / enums         apps/rtigo3/shaders/per_ray_data.h:39-71, function_indices.h:34-60,
 *                         light_definition.h:34-41, config.h:38-65
 *
 * Plain C11 / C++11 / CUDA; no CUDA or OptiX header is required so the CPU oracle (gcc) and
 * the device code (nvcc) see the same bytes.  Sizes and offsets are static-asserted below
 * against the values measured on the reference headers.
 *
 * Two slots change MEANING (not size or offset) because there is no OptiX and no texture unit
 * dependence in this core:
 *   SystemData::topObject   (OptixTraversableHandle, 8 B)  -> device pointer to the core's
 *                           two-level BVH descriptor, returned by rtc_ias_build().
 *   SystemData::envTexture  (cudaTextureObject_t, 8 B)     -> device pointer to envWidth*envHeight
 *                           RGBA32F texels, row 0 = south pole.  The core filters it bilinearly in
 *                           software (wrap in u, clamp in v) so the CPU oracle can reproduce the
 *                           lookup bit-for-bit; the reference used the texture unit
 *                           (apps/rtigo3/shaders/miss.cu:90, light_sample.cu:147).
 */
#ifndef RTIGO3_ABI_H
#define RTIGO3_ABI_H

#include <stddef.h>
#include <stdint.h>

#define RT_ALIGNAS(n) __attribute__((aligned(n)))   /* gcc, g++ and nvcc all accept this spelling on a struct tag */
#ifdef __cplusplus
#define RT_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define RT_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

/* ---- vector PODs, layout-identical to CUDA's float2/float3/float4/int2/int4 ---- */
typedef struct RT_ALIGNAS(8)  { float x, y; }        rt_float2;
typedef struct                { float x, y, z; }     rt_float3;
typedef struct RT_ALIGNAS(16) { float x, y, z, w; }  rt_float4;
typedef struct RT_ALIGNAS(8)  { int x, y; }          rt_int2;
typedef struct RT_ALIGNAS(16) { int x, y, z, w; }    rt_int4;

/* ---- constants (config.h:38-52, per_ray_data.h:39-71) ---- */
#define RT_DEFAULT_MAX            1.e27f
#define RT_SCENE_EPSILON_SCALE    1.0e-7f
#define RT_CLOCK_FACTOR_SCALE     1.0e-9f
#define RT_DENOMINATOR_EPSILON    1.0e-6f
#define RT_MICROFACET_MIN_ROUGHNESS 0.0014142f

#define RT_MATERIAL_STACK_EMPTY  (-1)
#define RT_MATERIAL_STACK_FIRST    0
#define RT_MATERIAL_STACK_LAST     3
#define RT_MATERIAL_STACK_SIZE     4

#define RT_FLAG_HIT           0x00000001u
#define RT_FLAG_SHADOW        0x00000002u
#define RT_FLAG_DIFFUSE       0x00000004u
#define RT_FLAG_FRONTFACE     0x00000010u
#define RT_FLAG_THINWALLED    0x00000020u
#define RT_FLAG_TRANSMISSION  0x00000100u
#define RT_FLAG_VOLUME        0x00001000u
#define RT_FLAG_TERMINATE     0x80000000u
#define RT_FLAG_CLEAR_MASK    RT_FLAG_DIFFUSE

#define RT_PI_F      3.14159265358979323846f
#define RT_1_PI_F    0.318309886183790671538f

/* ---- enums (function_indices.h, light_definition.h, Device.h:58-65) ---- */
enum { RT_RAYTYPE_RADIANCE = 0, RT_RAYTYPE_SHADOW = 1, RT_NUM_RAYTYPES = 2 };
enum { RT_LENS_PINHOLE = 0, RT_LENS_FISHEYE = 1, RT_LENS_SPHERE = 2, RT_NUM_LENS_SHADERS = 3 };
enum { RT_BRDF_DIFFUSE = 0, RT_BRDF_SPECULAR = 1, RT_BSDF_SPECULAR = 2, RT_BRDF_GGX_SMITH = 3,
       RT_BSDF_GGX_SMITH = 4, RT_NUM_BSDF_INDICES = 5 };
enum { RT_LIGHT_ENVIRONMENT = 0, RT_LIGHT_PARALLELOGRAM = 1, RT_NUM_LIGHT_TYPES = 2 };
enum { RT_MISS_NULL = 0, RT_MISS_CONSTANT = 1, RT_MISS_SPHERE = 2 };
enum { RT_STRATEGY_SINGLE_GPU = 0, RT_STRATEGY_MULTI_GPU_ZERO_COPY = 1,
       RT_STRATEGY_MULTI_GPU_PEER_ACCESS = 2, RT_STRATEGY_MULTI_GPU_LOCAL_COPY = 3,
       RT_NUM_STRATEGIES = 4 };

/* ---- per-vertex record, stride 48 B: the BVH builder reads .vertex at offset 0 ---- */
typedef struct {
  rt_float3 vertex;
  rt_float3 tangent;
  rt_float3 normal;
  rt_float3 texcoord;
} rt_TriangleAttributes;

typedef struct {
  rt_float3 P, U, V, W;
} rt_CameraDefinition;

typedef struct {
  int       type;       /* RT_LIGHT_* */
  rt_float3 position;
  rt_float3 vecU;
  rt_float3 vecV;
  rt_float3 normal;
  float     area;
  rt_float3 emission;
  float     unused0, unused1, unused2;
} rt_LightDefinition;

typedef struct {
  uint64_t  textureAlbedo;  /* 0 = none (textures are a "next" row) */
  uint64_t  textureCutout;  /* 0 = none */
  rt_float2 roughness;
  int       indexBSDF;      /* RT_BRDF_* / RT_BSDF_* */
  rt_float3 albedo;
  rt_float3 absorption;
  float     ior;
  unsigned int flags;       /* RT_FLAG_THINWALLED or 0 */
  int       pad0;
} rt_MaterialDefinition;

typedef struct {
  uint64_t attributes;      /* device pointer to rt_TriangleAttributes[] */
  uint64_t indices;         /* device pointer to uint32 triplets */
  int      materialIndex;
  int      lightIndex;      /* negative: not a light */
} rt_GeometryInstanceData;

typedef struct {
  rt_int4  rect;
  uint64_t topObject;            /* see header comment */
  uint64_t outputBuffer;         /* float4[resolution.y][resolution.x], row 0 = bottom */
  uint64_t tileBuffer;
  uint64_t texelBuffer;          /* float4[resolution.y][launchWidth] (local-copy strategy) */
  uint64_t cameraDefinitions;    /* rt_CameraDefinition* */
  uint64_t lightDefinitions;     /* rt_LightDefinition*  */
  uint64_t materialDefinitions;  /* rt_MaterialDefinition* */
  uint64_t envTexture;           /* see header comment */
  uint64_t envCDF_U;             /* float[(envWidth+1)*envHeight] */
  uint64_t envCDF_V;             /* float[envHeight+1] */
  rt_int2  resolution;
  rt_int2  tileSize;
  rt_int2  tileShift;
  rt_int2  pathLengths;          /* .x = min length before Russian roulette, .y = max length */
  int      deviceCount;
  int      deviceIndex;
  int      distribution;
  int      iterationIndex;
  int      samplesSqrt;
  float    sceneEpsilon;
  float    clockScale;
  int      lensShader;
  int      numCameras;
  int      numMaterials;
  int      numLights;
  unsigned int envWidth;
  unsigned int envHeight;
  float    envIntegral;
  float    envRotation;
} rt_SystemData;

typedef struct {
  uint64_t outputBuffer;
  uint64_t tileBuffer;
  rt_int2  resolution;
  rt_int2  tileSize;
  rt_int2  tileShift;
  int      launchWidth;
  int      deviceCount;
  int      deviceIndex;
} rt_CompositorData;

/* Tonemapper parameters as the system description file names them
 * (apps/rtigo3/inc/TonemapperGUI.h; consumed by Application::screenshot, Application.cpp:2262-2295). */
typedef struct {
  float gamma;
  float colorBalance[3];
  float whitePoint;
  float burnHighlights;
  float crushBlacks;
  float saturation;
  float brightness;
} rt_TonemapperParams;

/* ---- layout pins ---- */
RT_STATIC_ASSERT(sizeof(rt_TriangleAttributes) == 48, "TriangleAttributes 48 B");
RT_STATIC_ASSERT(sizeof(rt_CameraDefinition) == 48, "CameraDefinition 48 B");
RT_STATIC_ASSERT(sizeof(rt_LightDefinition) == 80, "LightDefinition 80 B");
RT_STATIC_ASSERT(offsetof(rt_LightDefinition, position) == 4, "light.position");
RT_STATIC_ASSERT(offsetof(rt_LightDefinition, normal) == 40, "light.normal");
RT_STATIC_ASSERT(offsetof(rt_LightDefinition, area) == 52, "light.area");
RT_STATIC_ASSERT(offsetof(rt_LightDefinition, emission) == 56, "light.emission");
RT_STATIC_ASSERT(sizeof(rt_MaterialDefinition) == 64, "MaterialDefinition 64 B");
RT_STATIC_ASSERT(offsetof(rt_MaterialDefinition, roughness) == 16, "material.roughness");
RT_STATIC_ASSERT(offsetof(rt_MaterialDefinition, indexBSDF) == 24, "material.indexBSDF");
RT_STATIC_ASSERT(offsetof(rt_MaterialDefinition, albedo) == 28, "material.albedo");
RT_STATIC_ASSERT(offsetof(rt_MaterialDefinition, absorption) == 40, "material.absorption");
RT_STATIC_ASSERT(offsetof(rt_MaterialDefinition, ior) == 52, "material.ior");
RT_STATIC_ASSERT(offsetof(rt_MaterialDefinition, flags) == 56, "material.flags");
RT_STATIC_ASSERT(sizeof(rt_GeometryInstanceData) == 24, "GeometryInstanceData 24 B");
RT_STATIC_ASSERT(sizeof(rt_SystemData) == 192, "SystemData 192 B");
RT_STATIC_ASSERT(offsetof(rt_SystemData, topObject) == 16, "sys.topObject");
RT_STATIC_ASSERT(offsetof(rt_SystemData, outputBuffer) == 24, "sys.outputBuffer");
RT_STATIC_ASSERT(offsetof(rt_SystemData, texelBuffer) == 40, "sys.texelBuffer");
RT_STATIC_ASSERT(offsetof(rt_SystemData, cameraDefinitions) == 48, "sys.cameraDefinitions");
RT_STATIC_ASSERT(offsetof(rt_SystemData, envTexture) == 72, "sys.envTexture");
RT_STATIC_ASSERT(offsetof(rt_SystemData, envCDF_V) == 88, "sys.envCDF_V");
RT_STATIC_ASSERT(offsetof(rt_SystemData, resolution) == 96, "sys.resolution");
RT_STATIC_ASSERT(offsetof(rt_SystemData, pathLengths) == 120, "sys.pathLengths");
RT_STATIC_ASSERT(offsetof(rt_SystemData, deviceCount) == 128, "sys.deviceCount");
RT_STATIC_ASSERT(offsetof(rt_SystemData, iterationIndex) == 140, "sys.iterationIndex");
RT_STATIC_ASSERT(offsetof(rt_SystemData, sceneEpsilon) == 148, "sys.sceneEpsilon");
RT_STATIC_ASSERT(offsetof(rt_SystemData, lensShader) == 156, "sys.lensShader");
RT_STATIC_ASSERT(offsetof(rt_SystemData, numLights) == 168, "sys.numLights");
RT_STATIC_ASSERT(offsetof(rt_SystemData, envWidth) == 172, "sys.envWidth");
RT_STATIC_ASSERT(offsetof(rt_SystemData, envIntegral) == 180, "sys.envIntegral");
RT_STATIC_ASSERT(offsetof(rt_SystemData, envRotation) == 184, "sys.envRotation");
RT_STATIC_ASSERT(sizeof(rt_CompositorData) == 56, "CompositorData 56 B");

#endif /* RTIGO3_ABI_H */
