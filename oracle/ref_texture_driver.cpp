// oracle/ref_texture_driver.cpp -- TEST INFRASTRUCTURE (container only; reads /root/reference at BUILD time, never copied).
//
// Runs the reference's own Texture::calculateSphericalCDF (apps/rtigo3/src/Texture.cpp:1540-1645, compiled unmodified where it
// lies: `make -C oracle reftex`) on the host.  It is the producer of envCDF_U / envCDF_V / envIntegral which the spherical
// environment light and miss program read (light_sample.cu:77-144, miss.cu:75-109); host/EnvMap.cpp restates it and
// tests/golden/reference_envcdf.npz pins that restatement against this build.
//
// The function's only outside contacts are cuMemAlloc / cuMemcpyHtoD (it uploads the tables); they are served from host
// memory here.  DevIL is replaced by ref_shim/IL/il.h (enumerators only) and the Picture methods Texture::create() would call
// are never reached.
#define private public      // calculateSphericalCDF and the members it fills are private
#include "inc/Texture.h"
#undef private

#include <cstdlib>
#include <cstring>

// ---- CUDA driver API stand-ins (host memory) ---------------------------------------------------------------------------
extern "C" {
CUresult cuMemAlloc_v2(CUdeviceptr* dptr, size_t bytes) { *dptr = (CUdeviceptr)(uintptr_t)std::malloc(bytes ? bytes : 1); return *dptr ? CUDA_SUCCESS : CUDA_ERROR_OUT_OF_MEMORY; }
CUresult cuMemFree_v2(CUdeviceptr dptr) { std::free((void*)(uintptr_t)dptr); return CUDA_SUCCESS; }
CUresult cuMemcpyHtoD_v2(CUdeviceptr dst, const void* src, size_t bytes) { std::memcpy((void*)(uintptr_t)dst, src, bytes); return CUDA_SUCCESS; }
CUresult cuGetErrorName(CUresult, const char** s) { *s = "CUDA_ERROR (host stand-in)"; return CUDA_SUCCESS; }
CUresult cuGetErrorString(CUresult, const char** s) { *s = "host stand-in"; return CUDA_SUCCESS; }
// never reached by calculateSphericalCDF
CUresult cuArray3DCreate_v2(CUarray*, const CUDA_ARRAY3D_DESCRIPTOR*) { return CUDA_ERROR_NOT_SUPPORTED; }
CUresult cuArrayDestroy(CUarray) { return CUDA_SUCCESS; }
CUresult cuMemcpy3D_v2(const CUDA_MEMCPY3D*) { return CUDA_ERROR_NOT_SUPPORTED; }
CUresult cuMipmappedArrayCreate(CUmipmappedArray*, const CUDA_ARRAY3D_DESCRIPTOR*, unsigned int) { return CUDA_ERROR_NOT_SUPPORTED; }
CUresult cuMipmappedArrayDestroy(CUmipmappedArray) { return CUDA_SUCCESS; }
CUresult cuMipmappedArrayGetLevel(CUarray*, CUmipmappedArray, unsigned int) { return CUDA_ERROR_NOT_SUPPORTED; }
CUresult cuTexObjectCreate(CUtexObject*, const CUDA_RESOURCE_DESC*, const CUDA_TEXTURE_DESC*, const CUDA_RESOURCE_VIEW_DESC*) { return CUDA_ERROR_NOT_SUPPORTED; }
CUresult cuTexObjectDestroy(CUtexObject) { return CUDA_SUCCESS; }
}

// Picture methods referenced by Texture::create / update (not reached)
const Image* Picture::getImageLevel(unsigned int, unsigned int) const { return nullptr; }
unsigned int Picture::getNumberOfLevels(unsigned int) const { return 0; }
bool Picture::isCubemap() const { return false; }

// rgba: width * height RGBA32F texels, row 0 first.  cdfU: (width + 1) * height floats, cdfV: height + 1 floats.
extern "C" int reftex_spherical_cdf(const float* rgba, unsigned int width, unsigned int height, float* cdfU, float* cdfV, float* integral)
{
  Texture t;
  t.m_width = width; t.m_height = height; t.m_depth = 1;
  t.calculateSphericalCDF(rgba);
  std::memcpy(cdfU, (const void*)(uintptr_t)t.m_d_envCDF_U, sizeof(float) * (size_t)(width + 1) * height);
  std::memcpy(cdfV, (const void*)(uintptr_t)t.m_d_envCDF_V, sizeof(float) * (size_t)(height + 1));
  *integral = t.m_integral;
  return 0;
}
