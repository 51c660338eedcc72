/*
 * optix.h (shim) -- TEST INFRASTRUCTURE ONLY.
 *
 * Lets the reference's own device programs (/root/reference/apps/rtigo3/shaders/*.cu, never copied into this
 * repository) compile UNMODIFIED with g++ for the host, so that the scalar oracle can be checked against the
 * reference's arithmetic itself.  Only what those ten translation units use is declared: the launch/payload/hit
 * query intrinsics, optixTrace, optixDirectCall, tex2D.  The definitions live in oracle/ref_driver.cpp; traversal
 * (which OptiX keeps in the driver / RT cores and the reference has no source for) is served by the oracle's
 * intersector, so this pins everything EXCEPT the ray/triangle arithmetic.
 */
#ifndef REF_SHIM_OPTIX_H
#define REF_SHIM_OPTIX_H

#include <cuda_runtime.h>

typedef unsigned long long OptixTraversableHandle;
typedef unsigned long long CUdeviceptr_shim;
#ifndef CUDA_VERSION
typedef unsigned long long CUdeviceptr;
#endif
typedef unsigned int OptixVisibilityMask;

enum OptixRayFlags
{
  OPTIX_RAY_FLAG_NONE = 0u,
  OPTIX_RAY_FLAG_DISABLE_ANYHIT = 1u << 0,
  OPTIX_RAY_FLAG_ENFORCE_ANYHIT = 1u << 1,
  OPTIX_RAY_FLAG_TERMINATE_ON_FIRST_HIT = 1u << 2,
  OPTIX_RAY_FLAG_DISABLE_CLOSESTHIT = 1u << 3
};

uint3 optixGetLaunchIndex();
uint3 optixGetLaunchDimensions();
unsigned int optixGetPayload_0();
unsigned int optixGetPayload_1();
CUdeviceptr optixGetSbtDataPointer();
unsigned int optixGetPrimitiveIndex();
float2 optixGetTriangleBarycentrics();
float optixGetRayTmax();
OptixTraversableHandle optixGetTransformListHandle(unsigned int index);
const float4* optixGetInstanceTransformFromHandle(OptixTraversableHandle handle);
const float4* optixGetInstanceInverseTransformFromHandle(OptixTraversableHandle handle);
void optixTerminateRay();
void optixIgnoreIntersection();
unsigned int optixGetExceptionCode();

void ref_trace(OptixTraversableHandle handle, float3 origin, float3 direction, float tmin, float tmax, float rayTime,
               unsigned int visibilityMask, unsigned int rayFlags, unsigned int sbtOffset, unsigned int sbtStride,
               unsigned int missSbtIndex, unsigned int& p0, unsigned int& p1);

static inline void optixTrace(OptixTraversableHandle handle, float3 origin, float3 direction, float tmin, float tmax, float rayTime,
                              OptixVisibilityMask visibilityMask, unsigned int rayFlags, unsigned int sbtOffset, unsigned int sbtStride,
                              unsigned int missSbtIndex, unsigned int& p0, unsigned int& p1)
{
  ref_trace(handle, origin, direction, tmin, tmax, rayTime, visibilityMask, rayFlags, sbtOffset, sbtStride, missSbtIndex, p0, p1);
}

void* ref_callable(unsigned int sbtIndex);

template <typename ReturnT, typename... ArgTypes>
static inline ReturnT optixDirectCall(unsigned int sbtIndex, ArgTypes... args)
{
  typedef ReturnT (*Fn)(ArgTypes...);
  return reinterpret_cast<Fn>(ref_callable(sbtIndex))(args...);
}

template <typename T> T tex2D(cudaTextureObject_t texture, float u, float v);
template <> float4 tex2D<float4>(cudaTextureObject_t texture, float u, float v);

#endif
