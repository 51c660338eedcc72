// HostTypes.h -- host-side names for the device-layout structs (include/rtigo3_abi.h) and a few float3 helpers.
// The names follow apps/rtigo3/shaders/*.h so the host classes read like the reference's.
#pragma once
#include <cmath>
#include <string>

#include "rtigo3_abi.h"

using float2 = rt_float2;
using float3 = rt_float3;
using float4 = rt_float4;
using int2   = rt_int2;
using TriangleAttributes   = rt_TriangleAttributes;
using CameraDefinition     = rt_CameraDefinition;
using LightDefinition      = rt_LightDefinition;
using MaterialDefinition   = rt_MaterialDefinition;
using GeometryInstanceData = rt_GeometryInstanceData;
using SystemData           = rt_SystemData;
using CompositorData       = rt_CompositorData;
using TonemapperGUI        = rt_TonemapperParams;

enum RendererStrategy
{
  RS_INTERACTIVE_SINGLE_GPU = RT_STRATEGY_SINGLE_GPU,
  RS_INTERACTIVE_MULTI_GPU_ZERO_COPY = RT_STRATEGY_MULTI_GPU_ZERO_COPY,
  RS_INTERACTIVE_MULTI_GPU_PEER_ACCESS = RT_STRATEGY_MULTI_GPU_PEER_ACCESS,
  RS_INTERACTIVE_MULTI_GPU_LOCAL_COPY = RT_STRATEGY_MULTI_GPU_LOCAL_COPY,
  NUM_RENDERER_STRATEGIES = RT_NUM_STRATEGIES
};

enum LensShader { LENS_SHADER_PINHOLE = RT_LENS_PINHOLE, LENS_SHADER_FISHEYE = RT_LENS_FISHEYE, LENS_SHADER_SPHERE = RT_LENS_SPHERE };

enum FunctionIndex
{
  INDEX_BRDF_DIFFUSE = RT_BRDF_DIFFUSE, INDEX_BRDF_SPECULAR = RT_BRDF_SPECULAR, INDEX_BSDF_SPECULAR = RT_BSDF_SPECULAR,
  INDEX_BRDF_GGX_SMITH = RT_BRDF_GGX_SMITH, INDEX_BSDF_GGX_SMITH = RT_BSDF_GGX_SMITH, NUM_BSDF_INDICES = RT_NUM_BSDF_INDICES
};

// Host side GUI material parameters (apps/rtigo3/inc/MaterialGUI.h:39-51)
struct MaterialGUI
{
  std::string   name;
  FunctionIndex indexBSDF = INDEX_BRDF_DIFFUSE;
  float3        albedo = { 1.0f, 1.0f, 1.0f };
  float3        absorptionColor = { 1.0f, 1.0f, 1.0f };
  float         absorptionScale = 0.0f;
  float         ior = 1.5f;
  bool          thinwalled = false;
  bool          useAlbedoTexture = false;
  bool          useCutoutTexture = false;
  float2        roughness = { 0.1f, 0.1f };
};

// apps/rtigo3/inc/Device.h DeviceState
struct DeviceState
{
  int2       resolution = { 1, 1 };
  int2       tileSize = { 8, 8 };
  int2       pathLengths = { 0, 2 };
  int        distribution = 0;
  int        samplesSqrt = 1;
  LensShader lensShader = LENS_SHADER_PINHOLE;
  float      epsilonFactor = 500.0f;
  float      envRotation = 0.0f;
  float      clockFactor = 1000.0f;
};

inline float3 make_float3(float x, float y, float z) { float3 r; r.x = x; r.y = y; r.z = z; return r; }
inline float3 make_float3(float s) { return make_float3(s, s, s); }
inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
inline int2   make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }
inline float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator-(float3 a) { return make_float3(-a.x, -a.y, -a.z); }
inline float3 operator*(float3 a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
inline float3 operator*(float s, float3 a) { return make_float3(s * a.x, s * a.y, s * a.z); }
inline float3 operator/(float3 a, float s) { const float inv = 1.0f / s; return a * inv; }
inline float  dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(float3 a, float3 b) { return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float  length(float3 a) { return std::sqrt(dot(a, a)); }
inline float3 normalize(float3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }
inline bool operator!=(int2 a, int2 b) { return a.x != b.x || a.y != b.y; }
