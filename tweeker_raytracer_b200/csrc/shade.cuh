// shade.cuh -- device restatement of rtigo3's lens shaders, BSDF callables, light samplers, miss
// programs and the closest-hit shading logic.  Compiled with -fmad=false: every expression is evaluated
// with individually rounded IEEE operations in the order the reference writes them, and every
// transcendental comes from include/rt_portable_math.h, so radiance is reproducible bit for bit on the host.
//
// Reference (apps/rtigo3/shaders/): lens_shader.cu:40-99, bxdf_diffuse.cu:39-95, bxdf_specular.cu:42-134,
// bxdf_ggx_smith.cu:45-319, light_sample.cu:42-177, miss.cu:41-109, closesthit.cu:126-305,
// shader_common.h:47-187, random_number_generators.h:39-78, vector_math.h:547-608.
#pragma once

#include "rtc_internal.h"
#include "rt_portable_math.h"

#define SD __device__ __forceinline__

// ---- float3 algebra in the reference's evaluation order (vector_math.h) ----
SD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
SD float3 f3(float s) { return make_float3(s, s, s); }
SD float3 f3(const rt_float3& v) { return make_float3(v.x, v.y, v.z); }
SD float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
SD float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
SD float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
SD float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
SD float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
SD float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
SD float  dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SD float3 cross(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
SD float  length(float3 a) { return sqrtf(dot(a, a)); }
SD float3 normalize(float3 a) { const float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
SD float3 divs(float3 a, float s) { const float inv = 1.0f / s; return a * inv; }
SD float3 reflect(float3 i, float3 n) { return i - 2.0f * n * dot(n, i); }
SD bool   is_null(float3 v) { return v.x == 0.0f && v.y == 0.0f && v.z == 0.0f; }
SD float  fmax3(float3 a) { return fmaxf(fmaxf(a.x, a.y), a.z); }
SD float  intensity3(float3 c) { return (c.x + c.y + c.z) * 0.3333333333f; }
SD float  power_heuristic(float a, float b) { const float t = a * a; return t / (t + b * b); }

// ---- RNG (random_number_generators.h:39-78) ----
SD uint32_t tea4(uint32_t v0, uint32_t v1)
{
  uint32_t s0 = 0;
#pragma unroll
  for (int n = 0; n < 4; ++n)
  {
    s0 += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xA341316Cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xC8013EA4u);
    v1 += ((v0 << 4) + 0xAD90777Du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7E95761Eu);
  }
  return v0;
}
SD float rng(uint32_t& s) { s = s * 1664525u + 1013904223u; return (float)(s & 0x00FFFFFFu) / (float)0x01000000u; }
SD float2 rng2(uint32_t& s) { float2 r; r.x = rng(s); r.y = rng(s); return r; }

struct Prd
{
  float3 pos; float distance;
  float3 wo, wi;
  float3 radiance; uint32_t flags;
  float3 f_over_pdf; float pdf;
  float3 sigma_t; float2 ior;
  float4 absorption_ior;
  uint32_t seed;
};

struct State { float3 normalGeo, tangent, normal, albedo; };

struct Tbn
{
  float3 tangent, bitangent, normal;
  SD Tbn(float3 tangent_reference, float3 n) : normal(n)
  {
    bitangent = normalize(cross(normal, tangent_reference));
    tangent = cross(bitangent, normal);
  }
  SD float3 toLocal(float3 p) const { return f3(dot(p, tangent), dot(p, bitangent), dot(p, normal)); }
  SD float3 toWorld(float3 p) const { return p.x * tangent + p.y * bitangent + p.z * normal; }
};

SD bool refract_dir(float3& r, float3 i, float3 n, float ior)
{
  float3 nn = n;
  float negNdotV = dot(i, nn);
  float eta;
  if (negNdotV > 0.0f) { eta = ior; nn = -n; negNdotV = -negNdotV; }
  else                 { eta = 1.f / ior; }
  const float k = 1.f - eta * eta * (1.f - negNdotV * negNdotV);
  if (k < 0.0f) { r = f3(0.f); return false; }
  r = normalize(eta * i - (eta * negNdotV + sqrtf(k)) * nn);
  return true;
}

SD float fresnel_dielectric(float et, float cosIn)
{
  const float cosi = fabsf(cosIn);
  float sint = 1.0f - cosi * cosi;
  sint = (0.0f < sint) ? sqrtf(sint) / et : 0.0f;
  if (1.0f < sint) return 1.0f;
  float cost = 1.0f - sint * sint;
  cost = (0.0f < cost) ? sqrtf(cost) : 0.0f;
  const float et_cosi = et * cosi, et_cost = et * cost;
  const float rPerp = (cosi - et_cost) / (cosi + et_cost);
  const float rPar  = (et_cosi - cost) / (et_cosi + cost);
  const float result = (rPar * rPar + rPerp * rPerp) * 0.5f;
  return (result <= 1.0f) ? result : 1.0f;
}

// ---- brdf_diffuse ----
SD void align_vector(float3 axis, float3& w)
{
  const float s = copysignf(1.0f, axis.z);
  w.z *= s;
  const float3 h = f3(axis.x, axis.y, axis.z + s);
  const float k = dot(w, h) / (1.0f + fabsf(axis.z));
  w = k * h - w;
}

SD void sample_brdf_diffuse(const rt_MaterialDefinition& m, const State& st, Prd& prd)
{
  const float2 sample = rng2(prd.seed);
  const float theta = 2.0f * RT_PI_F * sample.x;
  const float r = sqrtf(sample.y);
  float3 w;
  w.x = r * rt_cosf(theta);
  w.y = r * rt_sinf(theta);
  w.z = 1.0f - w.x * w.x - w.y * w.y;
  w.z = (0.0f < w.z) ? sqrtf(w.z) : 0.0f;
  prd.pdf = w.z * RT_1_PI_F;
  align_vector(st.normal, w);
  prd.wi = w;
  if (prd.pdf <= 0.0f || dot(prd.wi, st.normalGeo) <= 0.0f) { prd.flags |= RT_FLAG_TERMINATE; return; }
  prd.f_over_pdf = st.albedo;
  prd.flags |= RT_FLAG_DIFFUSE;
}

SD float4 eval_brdf_diffuse(const State& st, float3 wiL)
{
  const float3 f = st.albedo * RT_1_PI_F;
  const float pdf = fmaxf(0.0f, dot(wiL, st.normal) * RT_1_PI_F);
  return make_float4(f.x, f.y, f.z, pdf);
}

// ---- brdf_specular / bsdf_specular ----
SD void sample_brdf_specular(const State& st, Prd& prd)
{
  prd.wi = reflect(-prd.wo, st.normal);
  if (dot(prd.wi, st.normalGeo) <= 0.0f) { prd.flags |= RT_FLAG_TERMINATE; return; }
  prd.f_over_pdf = st.albedo;
  prd.pdf = 1.0f;
}

SD void sample_bsdf_specular(const rt_MaterialDefinition& m, const State& st, Prd& prd)
{
  prd.absorption_ior = make_float4(m.absorption.x, m.absorption.y, m.absorption.z, m.ior);
  const float eta = (prd.flags & (RT_FLAG_FRONTFACE | RT_FLAG_THINWALLED))
                  ? prd.absorption_ior.w / prd.ior.x
                  : prd.ior.y / prd.absorption_ior.w;
  const float3 R = reflect(-prd.wo, st.normal);
  float reflective = 1.0f;
  if (refract_dir(prd.wi, -prd.wo, st.normal, eta))
  {
    if (prd.flags & RT_FLAG_THINWALLED) prd.wi = -prd.wo;
    reflective = fresnel_dielectric(eta, dot(prd.wo, st.normal));
  }
  const float pseudo = rng(prd.seed);
  if (pseudo < reflective) prd.wi = R;
  else if (!(prd.flags & RT_FLAG_THINWALLED)) prd.flags |= RT_FLAG_TRANSMISSION;
  prd.f_over_pdf = st.albedo;
  prd.pdf = 1.0f;
}

// ---- GGX-Smith ----
SD float2 ggx_d_pdf(float ax, float ay, float3 wm)
{
  if (RT_DENOMINATOR_EPSILON < wm.z)
  {
    const float cosThetaSqr = wm.z * wm.z;
    const float tanThetaSqr = (1.0f - cosThetaSqr) / cosThetaSqr;
    const float phiM = rt_atan2f(wm.y, wm.x);
    const float cosPhiM = rt_cosf(phiM), sinPhiM = rt_sinf(phiM);
    const float term = 1.0f + tanThetaSqr * ((cosPhiM * cosPhiM) / (ax * ax) + (sinPhiM * sinPhiM) / (ay * ay));
    const float d = 1.0f / (RT_PI_F * ax * ay * cosThetaSqr * cosThetaSqr * term * term);
    return make_float2(d, d * wm.z);
  }
  return make_float2(0.0f, 0.0f);
}

SD float3 ggx_sample(float ax, float ay, float u1, float u2)
{
  const float theta = rt_atanf(ay * sqrtf(u1) / sqrtf(1.0f - u1));
  const float phi = 2.0f * RT_PI_F * u2;
  const float sinTheta = rt_sinf(theta);
  return normalize(f3(rt_cosf(phi) * sinTheta * ax / ay, rt_sinf(phi) * sinTheta, rt_cosf(theta)));
}

SD float smith_g1(float alpha, float3 w, float3 wm)
{
  const float w_wm = dot(w, wm);
  if (w_wm * w.z <= 0.0f) return 0.0f;
  const float cosThetaSqr = w.z * w.z;
  const float sinThetaSqr = 1.0f - cosThetaSqr;
  const float tanThetaSqr = (0.0f < sinThetaSqr) ? sinThetaSqr / cosThetaSqr : 0.0f;
  const float invASqr = alpha * alpha * tanThetaSqr;
  return 2.0f / (1.0f + sqrtf(1.0f + invASqr));
}

SD float ggx_g(float ax, float ay, float3 wo, float3 wi, float3 wm)
{
  float phi = rt_atan2f(wo.y, wo.x);
  float c = rt_cosf(phi), s = rt_sinf(phi);
  float alpha = sqrtf(c * c * ax * ax + s * s * ay * ay);
  const float g = smith_g1(alpha, wo, wm);
  phi = rt_atan2f(wi.y, wi.x);
  c = rt_cosf(phi); s = rt_sinf(phi);
  alpha = sqrtf(c * c * ax * ax + s * s * ay * ay);
  return g * smith_g1(alpha, wi, wm);
}

SD void sample_brdf_ggx(const rt_MaterialDefinition& m, const State& st, Prd& prd)
{
  const float2 sample = rng2(prd.seed);
  const float3 wm = ggx_sample(m.roughness.x, m.roughness.y, sample.x, sample.y);
  const Tbn ts(st.tangent, st.normal);
  const float3 wh = ts.toWorld(wm);
  prd.wi = reflect(-prd.wo, wh);
  if (dot(prd.wi, st.normalGeo) <= 0.0f) { prd.flags |= RT_FLAG_TERMINATE; return; }
  const float3 wo = ts.toLocal(prd.wo);
  const float3 wi = ts.toLocal(prd.wi);
  const float wi_wh = dot(prd.wi, wh);
  if (wo.z <= 0.0f || wi.z <= 0.0f || wi_wh <= 0.0f) { prd.flags |= RT_FLAG_TERMINATE; return; }
  const float2 D_PDF = ggx_d_pdf(m.roughness.x, m.roughness.y, wm);
  if (D_PDF.y <= 0.0f) { prd.flags |= RT_FLAG_TERMINATE; return; }
  const float G = ggx_g(m.roughness.x, m.roughness.y, wo, wi, wm);
  prd.pdf = D_PDF.y / (4.0f * wi_wh);
  prd.f_over_pdf = st.albedo * (G * D_PDF.x * wi_wh / (D_PDF.y * wo.z));
  prd.flags |= RT_FLAG_DIFFUSE;
}

SD float4 eval_brdf_ggx(const rt_MaterialDefinition& m, const State& st, const Prd& prd, float3 wiL)
{
  const Tbn ts(st.tangent, st.normal);
  const float3 wo = ts.toLocal(prd.wo);
  const float3 wi = ts.toLocal(wiL);
  if (wo.z <= 0.0f || wi.z <= 0.0f) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  float3 wm = wo + wi;
  if (is_null(wm)) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  wm = normalize(wm);
  const float2 D_PDF = ggx_d_pdf(m.roughness.x, m.roughness.y, wm);
  const float G = ggx_g(m.roughness.x, m.roughness.y, wo, wi, wm);
  const float3 f = st.albedo * (D_PDF.x * G / (4.0f * wo.z * wi.z));
  const float pdf = D_PDF.y / (4.0f * dot(wi, wm));
  return make_float4(f.x, f.y, f.z, pdf);
}

SD void sample_bsdf_ggx(const rt_MaterialDefinition& m, const State& st, Prd& prd)
{
  prd.absorption_ior = make_float4(m.absorption.x, m.absorption.y, m.absorption.z, m.ior);
  const float eta = (prd.flags & (RT_FLAG_FRONTFACE | RT_FLAG_THINWALLED))
                  ? prd.absorption_ior.w / prd.ior.x
                  : prd.ior.y / prd.absorption_ior.w;
  const float2 sample = rng2(prd.seed);
  const float3 wm = ggx_sample(m.roughness.x, m.roughness.y, sample.x, sample.y);
  const Tbn ts(st.tangent, st.normal);
  const float3 wh = ts.toWorld(wm);
  const float3 R = reflect(-prd.wo, wh);
  float reflective = 1.0f;
  if (refract_dir(prd.wi, -prd.wo, wh, eta))
  {
    if (prd.flags & RT_FLAG_THINWALLED) prd.wi = reflect(R, st.normal);
    reflective = fresnel_dielectric(eta, dot(prd.wo, wh));
  }
  const float pseudo = rng(prd.seed);
  if (pseudo < reflective) prd.wi = R;
  else if (!(prd.flags & RT_FLAG_THINWALLED)) prd.flags |= RT_FLAG_TRANSMISSION;
  prd.f_over_pdf = st.albedo;
  prd.pdf = 1.0f;
}

// The direct-callable table of the reference (closesthit.cu:246-248) as a switch.
SD void bsdf_sample(const rt_MaterialDefinition& m, const State& st, Prd& prd)
{
  switch (m.indexBSDF)
  {
    default:
    case RT_BRDF_DIFFUSE:   sample_brdf_diffuse(m, st, prd); break;
    case RT_BRDF_SPECULAR:  sample_brdf_specular(st, prd); break;
    case RT_BSDF_SPECULAR:  sample_bsdf_specular(m, st, prd); break;
    case RT_BRDF_GGX_SMITH: sample_brdf_ggx(m, st, prd); break;
    case RT_BSDF_GGX_SMITH: sample_bsdf_ggx(m, st, prd); break;
  }
}

SD float4 bsdf_eval(const rt_MaterialDefinition& m, const State& st, const Prd& prd, float3 wiL)
{
  switch (m.indexBSDF)
  {
    case RT_BRDF_DIFFUSE:   return eval_brdf_diffuse(st, wiL);
    case RT_BRDF_GGX_SMITH: return eval_brdf_ggx(m, st, prd, wiL);
    default:                return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  }
}

// ---- environment texture: software bilinear filter, wrap u / clamp v (replaces tex2D, miss.cu:90, light_sample.cu:147) ----
SD float3 env_lookup(const rt_SystemData& sys, float u, float v)
{
  const int W = (int)sys.envWidth, H = (int)sys.envHeight;
  const float4* tex = reinterpret_cast<const float4*>(sys.envTexture);
  const float x = u * (float)W - 0.5f, y = v * (float)H - 0.5f;
  const float fx = floorf(x), fy = floorf(y);
  const float ax = x - fx, ay = y - fy;
  int x0 = (int)fx % W; if (x0 < 0) x0 += W;
  int x1 = x0 + 1; if (x1 >= W) x1 = 0;
  int y0 = (int)fy, y1 = y0 + 1;
  if (y0 < 0) y0 = 0; if (y0 > H - 1) y0 = H - 1;
  if (y1 < 0) y1 = 0; if (y1 > H - 1) y1 = H - 1;
  const float4 t00 = __ldg(tex + (size_t)y0 * W + x0), t10 = __ldg(tex + (size_t)y0 * W + x1);
  const float4 t01 = __ldg(tex + (size_t)y1 * W + x0), t11 = __ldg(tex + (size_t)y1 * W + x1);
  const float3 a = f3(t00.x + ax * (t10.x - t00.x), t00.y + ax * (t10.y - t00.y), t00.z + ax * (t10.z - t00.z));
  const float3 b = f3(t01.x + ax * (t11.x - t01.x), t01.y + ax * (t11.y - t01.y), t01.z + ax * (t11.z - t01.z));
  return f3(a.x + ay * (b.x - a.x), a.y + ay * (b.y - a.y), a.z + ay * (b.z - a.z));
}

// ---- material textures: software bilinear filter, wrap u AND v (replaces tex2D at closesthit.cu:235, anyhit.cu:70, :119;
// the reference's sampler is wrap/wrap, linear, normalised coordinates, src/Texture.cpp:670-675).  handle = address of a
// 16-byte header {width, height, 0, 0} followed by the RGBA32F texels (rtc_texture_create).
SD float3 tex2d_wrap(uint64_t handle, float u, float v)
{
  const uint4 header = __ldg(reinterpret_cast<const uint4*>(handle));
  const float4* tex = reinterpret_cast<const float4*>(handle) + 1;
  const int W = (int)header.x, H = (int)header.y;
  const float x = u * (float)W - 0.5f, y = v * (float)H - 0.5f;
  const float fx = floorf(x), fy = floorf(y);
  const float ax = x - fx, ay = y - fy;
  int x0 = (int)fx % W; if (x0 < 0) x0 += W;
  int x1 = x0 + 1; if (x1 >= W) x1 = 0;
  int y0 = (int)fy % H; if (y0 < 0) y0 += H;
  int y1 = y0 + 1; if (y1 >= H) y1 = 0;
  const float4 t00 = __ldg(tex + (size_t)y0 * W + x0), t10 = __ldg(tex + (size_t)y0 * W + x1);
  const float4 t01 = __ldg(tex + (size_t)y1 * W + x0), t11 = __ldg(tex + (size_t)y1 * W + x1);
  const float3 a = f3(t00.x + ax * (t10.x - t00.x), t00.y + ax * (t10.y - t00.y), t00.z + ax * (t10.z - t00.z));
  const float3 b = f3(t01.x + ax * (t11.x - t01.x), t01.y + ax * (t11.y - t01.y), t01.z + ax * (t11.z - t01.z));
  return f3(a.x + ay * (b.x - a.x), a.y + ay * (b.y - a.y), a.z + ay * (b.z - a.z));
}

// ---- lights ----
struct LightSample { float3 direction; float distance; float3 emission; float pdf; int index; };

SD void light_env_constant(int numLights, float2 sample, LightSample& ls)
{
  float3 p;
  p.z = 1.0f - 2.0f * sample.x;
  float r = 1.0f - p.z * p.z;
  r = (0.0f < r) ? sqrtf(r) : 0.0f;
  const float phi = sample.y * 2.0f * RT_PI_F;
  p.x = r * rt_cosf(phi);
  p.y = r * rt_sinf(phi);
  ls.direction = p;
  ls.pdf = 0.25f * RT_1_PI_F;
  ls.distance = RT_DEFAULT_MAX;
  ls.emission = f3((float)numLights);
}

SD void light_env_sphere(const rt_SystemData& sys, float2 sample, LightSample& ls)
{
  const unsigned int sizeV = sys.envHeight;
  unsigned int ilo = 0, ihi = sizeV;
  const float* cdfV = reinterpret_cast<const float*>(sys.envCDF_V);
  while (ilo != ihi - 1)
  {
    const unsigned int i = (ilo + ihi) >> 1;
    if (sample.y < __ldg(cdfV + i)) ihi = i; else ilo = i;
  }
  const unsigned int vIdx = ilo;
  const unsigned int sizeU = sys.envWidth;
  ilo = 0; ihi = sizeU;
  const float* cdfU = reinterpret_cast<const float*>(sys.envCDF_U) + (size_t)vIdx * (sizeU + 1);
  while (ilo != ihi - 1)
  {
    const unsigned int i = (ilo + ihi) >> 1;
    if (sample.x < __ldg(cdfU + i)) ihi = i; else ilo = i;
  }
  const unsigned int uIdx = ilo;
  const float cdfLowerU = __ldg(cdfU + uIdx), cdfUpperU = __ldg(cdfU + uIdx + 1);
  const float du = (sample.x - cdfLowerU) / (cdfUpperU - cdfLowerU);
  const float cdfLowerV = __ldg(cdfV + vIdx), cdfUpperV = __ldg(cdfV + vIdx + 1);
  const float dv = (sample.y - cdfLowerV) / (cdfUpperV - cdfLowerV);
  const float u = ((float)uIdx + du) / (float)sizeU;
  const float v = ((float)vIdx + dv) / (float)sizeV;
  const float phi = (u - sys.envRotation) * 2.0f * RT_PI_F;
  const float theta = v * RT_PI_F;
  const float sinTheta = rt_sinf(theta);
  ls.direction = f3(-rt_sinf(phi) * sinTheta, -rt_cosf(theta), rt_cosf(phi) * sinTheta);
  ls.distance = RT_DEFAULT_MAX;
  const float3 emission = env_lookup(sys, u, v);
  ls.emission = emission * (float)sys.numLights;
  ls.pdf = intensity3(emission) / sys.envIntegral;
}

SD void light_parallelogram(const rt_SystemData& sys, float3 point, float2 sample, LightSample& ls)
{
  ls.pdf = 0.0f;
  const rt_LightDefinition& light = reinterpret_cast<const rt_LightDefinition*>(sys.lightDefinitions)[ls.index];
  const float3 position = f3(light.position) + f3(light.vecU) * sample.x + f3(light.vecV) * sample.y;
  ls.direction = position - point;
  ls.distance = length(ls.direction);
  if (RT_DENOMINATOR_EPSILON < ls.distance)
  {
    ls.direction = divs(ls.direction, ls.distance);
    const float cosTheta = dot(-ls.direction, f3(light.normal));
    if (RT_DENOMINATOR_EPSILON < cosTheta)
    {
      ls.emission = f3(light.emission) * (float)sys.numLights;
      ls.pdf = (ls.distance * ls.distance) / (light.area * cosTheta);
    }
  }
}

// ---- lens shaders ----
SD void lens_shader(const rt_SystemData& sys, float2 screen, float2 pixel, float2 sample, float3& origin, float3& direction)
{
  const rt_CameraDefinition& cam = reinterpret_cast<const rt_CameraDefinition*>(sys.cameraDefinitions)[0];
  const float3 cU = f3(cam.U), cV = f3(cam.V), cW = f3(cam.W);
  origin = f3(cam.P);
  if (sys.lensShader == RT_LENS_FISHEYE)
  {
    const float2 fragment = make_float2(pixel.x + sample.x, pixel.y + sample.y);
    const float2 center = make_float2(screen.x * 0.5f, screen.y * 0.5f);
    const float invLen = 1.0f / sqrtf(center.x * center.x + center.y * center.y);
    const float2 uv = make_float2((fragment.x - center.x) * invLen, (fragment.y - center.y) * invLen);
    const float z = rt_cosf(sqrtf(uv.x * uv.x + uv.y * uv.y) * 0.7071067812f * 0.5f * RT_PI_F);
    const float3 U = normalize(cU), V = normalize(cV), W = normalize(cW);
    direction = normalize(uv.x * U + uv.y * V + z * W);
  }
  else if (sys.lensShader == RT_LENS_SPHERE)
  {
    const float2 uv = make_float2((pixel.x + sample.x) / screen.x, (pixel.y + sample.y) / screen.y);
    const float phi = uv.x * 2.0f * RT_PI_F;
    const float theta = uv.y * RT_PI_F;
    const float sinTheta = rt_sinf(theta);
    const float3 v = f3(-rt_sinf(phi) * sinTheta, -rt_cosf(theta), -rt_cosf(phi) * sinTheta);
    const float3 U = normalize(cU), V = normalize(cV), W = normalize(cW);
    direction = normalize(v.x * U + v.y * V + v.z * W);
  }
  else
  {
    const float2 fragment = make_float2(pixel.x + sample.x, pixel.y + sample.y);
    const float2 ndc = make_float2((fragment.x / screen.x) * 2.0f - 1.0f, (fragment.y / screen.y) * 2.0f - 1.0f);
    direction = normalize(cU * ndc.x + cV * ndc.y + cW);
  }
}

// ---- miss programs ----
SD void miss_program(const rt_SystemData& sys, int miss, Prd& prd)
{
  if (miss == RT_MISS_CONSTANT)
  {
    const float w = (prd.flags & RT_FLAG_DIFFUSE) ? power_heuristic(prd.pdf, 0.25f * RT_1_PI_F) : 1.0f;
    prd.radiance = f3(w);
  }
  else if (miss == RT_MISS_SPHERE)
  {
    const float3 R = prd.wi;
    const float u = (rt_atan2f(R.x, -R.z) + RT_PI_F) * 0.5f * RT_1_PI_F + sys.envRotation;
    const float theta = rt_acosf(-R.y);
    const float v = theta * RT_1_PI_F;
    const float3 emission = env_lookup(sys, u, v);
    float w = 1.0f;
    if (prd.flags & RT_FLAG_DIFFUSE)
    {
      const float pdfLight = intensity3(emission) / sys.envIntegral;
      w = power_heuristic(prd.pdf, pdfLight);
    }
    prd.radiance = emission * w;
  }
  else
  {
    prd.radiance = f3(0.0f);
  }
  prd.flags |= RT_FLAG_TERMINATE;
}

// raygeneration.cu:152-164
SD uint32_t distribute(const rt_SystemData& sys, uint32_t x, uint32_t y)
{
  const uint32_t xBlock = x >> sys.tileShift.x;
  const uint32_t yBlock = y >> sys.tileShift.y;
  const uint32_t xTile = xBlock * (uint32_t)sys.deviceCount + (((uint32_t)sys.deviceIndex + yBlock) % (uint32_t)sys.deviceCount);
  return xTile * (uint32_t)sys.tileSize.x + (x & (uint32_t)(sys.tileSize.x - 1));
}

// raygeneration.cu:173-201.  Returns false when the launch index falls outside the image.
SD bool start_path(const rt_SystemData& sys, uint32_t launchWidth, uint32_t x, uint32_t y, int iteration,
                   uint32_t& seed, float3& pos, float3& wi, uint32_t& column)
{
  uint32_t col = x;
  if (sys.distribution && 1 < sys.deviceCount)
  {
    col = distribute(sys, x, y);
    if ((uint32_t)sys.resolution.x <= col) return false;
  }
  const uint32_t seedIndex = launchWidth * y + col * (uint32_t)sys.deviceCount + (uint32_t)sys.deviceIndex;
  seed = tea4(seedIndex, (uint32_t)iteration);
  const float2 screen = make_float2((float)sys.resolution.x, (float)sys.resolution.y);
  const float2 pixel = make_float2((float)col, (float)y);
  const float2 sample = rng2(seed);
  lens_shader(sys, screen, pixel, sample, pos, wi);
  column = col;
  return true;
}
