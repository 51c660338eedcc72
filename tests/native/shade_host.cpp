// tests/native/shade_host.cpp -- TEST INFRASTRUCTURE: the PRODUCT's ray-generation source compiled for the host.
//
// csrc/shade.cuh is included unchanged (its only CUDA intrinsic is __ldg) and compiled by g++ with the flags that mirror the
// device build of the shading translation unit (no FMA contraction, IEEE division and square root): the TEA / LCG generators,
// distribute(), start_path() with the three lens shaders, and the software texture fetch run here as they do in
// k_extend_primary / k_generate_primary / the textured shade kernels.  tests/test_cpu_shade_source.py holds them against the
// oracle (which is pinned against the reference's own sources): primary rays bit for bit for every lens shader and every
// device of a tiled multi-GPU launch, the generators against the reference's golden vectors, texture fetches bit for bit.
#include <cmath>
#include <cstdint>
#include <cstring>

#include <cuda_runtime.h>

template <class T> static inline T __ldg(const T* p) { return *p; }

#include "shade.cuh"

extern "C" {

uint32_t sh_tea4(uint32_t v0, uint32_t v1) { return tea4(v0, v1); }

void sh_rng_sequence(uint32_t seed, int n, float* out, uint32_t* stateOut)
{
  for (int i = 0; i < n; ++i) out[i] = rng(seed);
  *stateOut = seed;
}

// k_generate_primary (csrc/kernels_shade.cu): one rtc_ray per launch index, tmax = -1 for skipped indices
void sh_generate_primary(const rt_SystemData* sys, uint32_t w, uint32_t h, int iteration, float* rays)
{
  for (uint32_t idx = 0; idx < w * h; ++idx)
  {
    const uint32_t y = idx / w, x = idx - y * w;
    uint32_t seed, col; float3 pos, wi;
    float* r = rays + 8 * (size_t)idx;
    if (start_path(*sys, w, x, y, iteration, seed, pos, wi, col))
    {
      r[0] = pos.x; r[1] = pos.y; r[2] = pos.z; r[3] = sys->sceneEpsilon;
      r[4] = wi.x; r[5] = wi.y; r[6] = wi.z; r[7] = RT_DEFAULT_MAX;
    }
    else
    {
      for (int k = 0; k < 7; ++k) r[k] = 0.0f;
      r[7] = -1.0f;
    }
  }
}

// tex2d_wrap: handle = address of {uint32 w, h, 0, 0} followed by w * h RGBA32F texels (rtc_texture_create)
void sh_tex2d(uint64_t handle, int n, const float* uv, float* rgb)
{
  for (int i = 0; i < n; ++i)
  {
    const float3 c = tex2d_wrap(handle, uv[2 * i], uv[2 * i + 1]);
    rgb[3 * i] = c.x; rgb[3 * i + 1] = c.y; rgb[3 * i + 2] = c.z;
  }
}

} // extern "C"
