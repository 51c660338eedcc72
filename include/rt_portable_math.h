/*
 * rt_portable_math.h -- the transcendental functions of the rtigo3 hot path, DEFINED.
 *
 * The reference device programs call sinf/cosf/atanf/atan2f/acosf/expf (bxdf_diffuse.cu:53-57,
 * bxdf_ggx_smith.cu:82-108, light_sample.cu:46-47,140-145, miss.cu:86-87, raygeneration.cu:97)
 * and are built with --use_fast_math (apps/rtigo3/CMakeLists.txt:181), so their last bits are
 * whatever the SFU approximations give on a particular GPU.  A CPU oracle cannot reproduce that,
 * and glibc's and libdevice's results differ from each other too.
 *
 * This core instead pins the arithmetic: every transcendental is a fixed sequence of IEEE-754
 * binary32 add/mul/div/sqrt/int-convert operations (Cody-Waite range reduction + a short minimax
 * polynomial, after the classic single-precision Cephes formulations).  Compiled with FMA
 * contraction OFF (nvcc -fmad=false, gcc -ffp-contract=off) the same source gives the same bits
 * on the GPU and on the host, which is what lets tests/ demand bit-exact radiance between the
 * sm_100a wavefront kernels and the scalar oracle.  Accuracy is 1-2 ulp on the ranges the path
 * tracer uses (checked against libm in tests/test_portable_math.py).
 *
 * Define RT_MATH_LIBM before including to route everything to <math.h> instead; the oracle uses
 * that build only to compare itself with the host-compiled reference shaders (oracle/_ref).
 */
#ifndef RT_PORTABLE_MATH_H
#define RT_PORTABLE_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD static inline
#endif

RT_HD uint32_t rt_float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
RT_HD float    rt_uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

#ifdef RT_MATH_LIBM

RT_HD float rt_sinf(float x) { return sinf(x); }
RT_HD float rt_cosf(float x) { return cosf(x); }
RT_HD float rt_atanf(float x) { return atanf(x); }
RT_HD float rt_atan2f(float y, float x) { return atan2f(y, x); }
RT_HD float rt_acosf(float x) { return acosf(x); }
RT_HD float rt_expf(float x) { return expf(x); }
RT_HD float rt_logf(float x) { return logf(x); }
RT_HD float rt_powf(float x, float y) { return powf(x, y); }

#else /* pinned arithmetic */

#define RT_M_FOPI   1.27323954473516f      /* 4/pi */
#define RT_M_DP1    0.78515625f            /* pi/4 split in three parts */
#define RT_M_DP2    2.4187564849853515625e-4f
#define RT_M_DP3    3.77489497744594108e-8f
#define RT_M_PIO2   1.5707963267948966192f
#define RT_M_PIO4   0.7853981633974483096f
#define RT_M_PI     3.14159265358979323846f

RT_HD float rt__sin_poly(float x, float z)
{
  float y = ((-1.9515295891E-4f * z + 8.3321608736E-3f) * z - 1.6666654611E-1f) * z * x;
  return y + x;
}

RT_HD float rt__cos_poly(float z)
{
  float y = ((2.443315711809948E-005f * z - 1.388731625493765E-003f) * z + 4.166664568298827E-002f) * z * z;
  y = y - 0.5f * z;
  return y + 1.0f;
}

RT_HD float rt_sinf(float xx)
{
  int   neg = (xx < 0.0f);
  float x   = fabsf(xx);
  int   j   = (int)(RT_M_FOPI * x);
  float y   = (float)j;
  if (j & 1) { j += 1; y += 1.0f; }
  j &= 7;
  if (j > 3) { neg = !neg; j -= 4; }
  x = ((x - y * RT_M_DP1) - y * RT_M_DP2) - y * RT_M_DP3;
  const float z = x * x;
  const float r = (j == 1 || j == 2) ? rt__cos_poly(z) : rt__sin_poly(x, z);
  return neg ? -r : r;
}

RT_HD float rt_cosf(float xx)
{
  int   neg = 0;
  float x   = fabsf(xx);
  int   j   = (int)(RT_M_FOPI * x);
  float y   = (float)j;
  if (j & 1) { j += 1; y += 1.0f; }
  j &= 7;
  if (j > 3) { j -= 4; neg = !neg; }
  if (j > 1) { neg = !neg; }
  x = ((x - y * RT_M_DP1) - y * RT_M_DP2) - y * RT_M_DP3;
  const float z = x * x;
  const float r = (j == 1 || j == 2) ? rt__sin_poly(x, z) : rt__cos_poly(z);
  return neg ? -r : r;
}

RT_HD float rt_atanf(float xx)
{
  const int neg = (xx < 0.0f);
  float x = fabsf(xx);
  float y;
  if (x > 2.414213562373095f)       { y = RT_M_PIO2; x = -(1.0f / x); }
  else if (x > 0.4142135623730950f) { y = RT_M_PIO4; x = (x - 1.0f) / (x + 1.0f); }
  else                              { y = 0.0f; }
  const float z = x * x;
  y += (((8.05374449538e-2f * z - 1.38776856032E-1f) * z + 1.99777106478E-1f) * z - 3.33329491539E-1f) * z * x + x;
  return neg ? -y : y;
}

RT_HD float rt_atan2f(float y, float x)
{
  if (x == 0.0f)
  {
    if (y > 0.0f) return RT_M_PIO2;
    if (y < 0.0f) return -RT_M_PIO2;
    return 0.0f;
  }
  if (y == 0.0f)
  {
    return (x > 0.0f) ? 0.0f : RT_M_PI;
  }
  float z = rt_atanf(y / x);
  if (x < 0.0f) z += (y < 0.0f) ? -RT_M_PI : RT_M_PI;
  return z;
}

RT_HD float rt__asinf_pos(float a) /* 0 <= a <= 1 */
{
  float x, z;
  int flag = 0;
  if (a < 1.0e-4f) return a;
  if (a > 0.5f) { z = 0.5f * (1.0f - a); x = sqrtf(z); flag = 1; }
  else          { x = a; z = x * x; }
  z = ((((4.2163199048E-2f * z + 2.4181311049E-2f) * z + 4.5470025998E-2f) * z + 7.4953002686E-2f) * z + 1.6666752422E-1f) * z * x + x;
  if (flag) { z = z + z; z = RT_M_PIO2 - z; }
  return z;
}

/* Arguments a rounding step outside [-1,1] (a normalised direction component) are clamped. */
RT_HD float rt_acosf(float x)
{
  if (x >  1.0f) x =  1.0f;
  if (x < -1.0f) x = -1.0f;
  if (x >  0.5f) return 2.0f * rt__asinf_pos(sqrtf(0.5f * (1.0f - x)));
  if (x < -0.5f) return RT_M_PI - 2.0f * rt__asinf_pos(sqrtf(0.5f * (1.0f + x)));
  const float s = rt__asinf_pos(fabsf(x));
  return RT_M_PIO2 - ((x < 0.0f) ? -s : s);
}

/* z * 2^n for z in [0.5, 2); exact except when the result is subnormal. */
RT_HD float rt__ldexpf(float z, int n)
{
  if (n > 127)  { z *= rt_uint_as_float((uint32_t)(127 + 127) << 23); n -= 127; if (n > 127) n = 127; }
  if (n < -126) { z *= rt_uint_as_float((uint32_t)(-100 + 127) << 23); n += 100; if (n < -126) n = -126; }
  return z * rt_uint_as_float((uint32_t)(n + 127) << 23);
}

RT_HD float rt_expf(float x)
{
  if (!(x == x)) return x;
  if (x > 88.72283905206835f) return rt_uint_as_float(0x7f800000u);
  if (x < -103.278929903431851103f) return 0.0f;
  float z = floorf(1.44269504088896341f * x + 0.5f);
  x = x - z * 0.693359375f;
  x = x - z * -2.12194440e-4f;
  const int n = (int)z;
  z = x * x;
  z = (((((1.9875691500E-4f * x + 1.3981999507E-3f) * x + 8.3334519073E-3f) * x + 4.1665795894E-2f) * x + 1.6666665459E-1f) * x + 5.0000001201E-1f) * z + x + 1.0f;
  return rt__ldexpf(z, n);
}

RT_HD float rt_logf(float xx)
{
  if (!(xx == xx)) return xx;
  if (xx < 0.0f) return rt_uint_as_float(0x7fc00000u);
  if (xx == 0.0f) return rt_uint_as_float(0xff800000u);
  if (xx == rt_uint_as_float(0x7f800000u)) return xx;
  int e = 0;
  uint32_t u = rt_float_as_uint(xx);
  if ((u & 0x7f800000u) == 0u) { xx *= 8388608.0f; e = -23; u = rt_float_as_uint(xx); } /* subnormal */
  e += (int)((u >> 23) & 0xffu) - 126;
  float x = rt_uint_as_float((u & 0x007fffffu) | 0x3f000000u); /* mantissa in [0.5, 1) */
  if (x < 0.707106781186547524f) { e -= 1; x = x + x - 1.0f; }
  else                           { x = x - 1.0f; }
  float z = x * x;
  float y = ((((((((7.0376836292E-2f * x - 1.1514610310E-1f) * x + 1.1676998740E-1f) * x - 1.2420140846E-1f) * x
              + 1.4249322787E-1f) * x - 1.6668057665E-1f) * x + 2.0000714765E-1f) * x - 2.4999993993E-1f) * x
              + 3.3333331174E-1f) * x * z;
  const float fe = (float)e;
  if (e) y += -2.12194440e-4f * fe;
  y += -0.5f * z;
  z = x + y;
  if (e) z += 0.693359375f * fe;
  return z;
}

/* Only used by the tonemapper on non-negative bases with positive exponents. */
RT_HD float rt_powf(float x, float y)
{
  if (x <= 0.0f) return 0.0f;
  return rt_expf(y * rt_logf(x));
}

#endif /* RT_MATH_LIBM */

#endif /* RT_PORTABLE_MATH_H */
