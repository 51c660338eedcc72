// tests/native/shade_host.cpp -- TEST INFRASTRUCTURE: the PRODUCT's ray-generation source compiled for the host.
//
// csrc/shade.cuh is included unchanged (its only CUDA intrinsic is __ldg) and compiled by g++ with the flags that mirror the
// device build of the shading translation unit (no FMA contraction, IEEE division and square root): the TEA / LCG generators,
// distribute(), start_path() with the three lens shaders, the software texture fetch, the five BSDF sample / eval callables, the
// three light callables and the three miss programs run here as they do in the kernels.  tests/test_cpu_shade_source.py holds them against the
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

// ---- BSDF and light callables, word for word against the oracle's test hooks (oracle/rt_oracle.c "TEST HOOKS") --------------
// prd = the 28 words of Prd in declaration order, st = normalGeo, tangent, normal, albedo.
static void prd_unpack(Prd& p, const uint32_t w[28])
{
  float f[28]; std::memcpy(f, w, sizeof(f));
  p.pos = f3(f[0], f[1], f[2]); p.distance = f[3];
  p.wo = f3(f[4], f[5], f[6]); p.wi = f3(f[7], f[8], f[9]);
  p.radiance = f3(f[10], f[11], f[12]); p.flags = w[13];
  p.f_over_pdf = f3(f[14], f[15], f[16]); p.pdf = f[17];
  p.sigma_t = f3(f[18], f[19], f[20]); p.ior = make_float2(f[21], f[22]);
  p.absorption_ior = make_float4(f[23], f[24], f[25], f[26]); p.seed = w[27];
}

static void prd_pack(const Prd& p, uint32_t w[28])
{
  const float f[28] = { p.pos.x, p.pos.y, p.pos.z, p.distance, p.wo.x, p.wo.y, p.wo.z, p.wi.x, p.wi.y, p.wi.z,
                        p.radiance.x, p.radiance.y, p.radiance.z, 0.0f, p.f_over_pdf.x, p.f_over_pdf.y, p.f_over_pdf.z, p.pdf,
                        p.sigma_t.x, p.sigma_t.y, p.sigma_t.z, p.ior.x, p.ior.y,
                        p.absorption_ior.x, p.absorption_ior.y, p.absorption_ior.z, p.absorption_ior.w, 0.0f };
  std::memcpy(w, f, sizeof(f));
  w[13] = p.flags; w[27] = p.seed;
}

static State state_unpack(const float st[12])
{
  State s;
  s.normalGeo = f3(st[0], st[1], st[2]); s.tangent = f3(st[3], st[4], st[5]); s.normal = f3(st[6], st[7], st[8]); s.albedo = f3(st[9], st[10], st[11]);
  return s;
}

void sh_bsdf_sample(const rt_MaterialDefinition* m, const float st[12], uint32_t prd[28])
{
  const State s = state_unpack(st);
  Prd p; prd_unpack(p, prd);
  bsdf_sample(*m, s, p);
  prd_pack(p, prd);
}

void sh_bsdf_eval(const rt_MaterialDefinition* m, const float st[12], const uint32_t prd[28], const float wiL[3], float out[4])
{
  const State s = state_unpack(st);
  Prd p; prd_unpack(p, prd);
  const float4 r = bsdf_eval(*m, s, p, f3(wiL[0], wiL[1], wiL[2]));
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

static void light_pack(const LightSample& ls, float out[8])
{
  out[0] = ls.direction.x; out[1] = ls.direction.y; out[2] = ls.direction.z; out[3] = ls.distance;
  out[4] = ls.emission.x; out[5] = ls.emission.y; out[6] = ls.emission.z; out[7] = ls.pdf;
}

void sh_light_constant(int numLights, const float sample[2], float out[8])
{
  LightSample ls; std::memset(&ls, 0, sizeof(ls));
  light_env_constant(numLights, make_float2(sample[0], sample[1]), ls);
  light_pack(ls, out);
}

void sh_light_parallelogram(const rt_LightDefinition* light, int numLights, const float point[3], const float sample[2], float out[8])
{
  rt_SystemData sys; std::memset(&sys, 0, sizeof(sys));
  sys.lightDefinitions = (uint64_t)(uintptr_t)light; sys.numLights = numLights;
  LightSample ls; std::memset(&ls, 0, sizeof(ls));
  ls.index = 0;
  light_parallelogram(sys, f3(point[0], point[1], point[2]), make_float2(sample[0], sample[1]), ls);
  light_pack(ls, out);
}

static rt_SystemData env_sys(const float* texels, uint32_t w, uint32_t h, const float* cdfU, const float* cdfV, float integral, float rotation, int numLights)
{
  rt_SystemData sys; std::memset(&sys, 0, sizeof(sys));
  sys.envTexture = (uint64_t)(uintptr_t)texels; sys.envCDF_U = (uint64_t)(uintptr_t)cdfU; sys.envCDF_V = (uint64_t)(uintptr_t)cdfV;
  sys.envWidth = w; sys.envHeight = h; sys.envIntegral = integral; sys.envRotation = rotation; sys.numLights = numLights;
  return sys;
}

void sh_miss(const float* texels, uint32_t w, uint32_t h, float integral, float rotation, int miss, uint32_t prd[28])
{
  const rt_SystemData sys = env_sys(texels, w, h, nullptr, nullptr, integral, rotation, 1);
  Prd p; prd_unpack(p, prd);
  miss_program(sys, miss, p);
  prd_pack(p, prd);
}

void sh_light_sphere(const float* texels, uint32_t w, uint32_t h, const float* cdfU, const float* cdfV, float integral, float rotation,
                     int numLights, const float sample[2], float out[8])
{
  const rt_SystemData sys = env_sys(texels, w, h, cdfU, cdfV, integral, rotation, numLights);
  LightSample ls; std::memset(&ls, 0, sizeof(ls));
  light_env_sphere(sys, make_float2(sample[0], sample[1]), ls);
  light_pack(ls, out);
}

} // extern "C"
