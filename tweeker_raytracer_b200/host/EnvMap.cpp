#include "EnvMap.h"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "ImageIO.h"
#include "rtigo3_abi.h"

void EnvMap::setTexels(unsigned int width, unsigned int height, const float* rgba)
{
  m_width = width; m_height = height;
  m_rgba.assign(rgba, rgba + (size_t)4 * width * height);
  calculateSphericalCDF();
}

// Analytic outdoor map: horizon-to-zenith gradient over a dark ground plus a small bright sun disc.
// Row y covers theta = pi * (y + 0.5) / height measured from the south pole, column x covers
// phi = 2 pi (x + 0.5) / width; the direction convention is miss.cu:86-87 (u from atan2(x, -z), v from acos(-y)).
bool EnvMap::createProcedural(unsigned int width, unsigned int height)
{
  if (width < 2 || height < 2) return false;
  std::vector<float> rgba((size_t)4 * width * height);
  const double sunU = 0.62, sunV = 0.72, sunRadius = 0.035;   // in (u,v) units
  for (unsigned int y = 0; y < height; ++y)
  {
    const double v = (y + 0.5) / height;             // 0 = straight down, 1 = straight up
    const double up = -std::cos(RT_PI_F * v);         // direction.y
    for (unsigned int x = 0; x < width; ++x)
    {
      const double u = (x + 0.5) / width;
      double r, g, b;
      if (up < 0.0) { const double k = 0.08 + 0.10 * (1.0 + up); r = 0.9 * k; g = 0.8 * k; b = 0.7 * k; }
      else
      {
        const double h = std::pow(1.0 - up, 3.0);
        r = 0.25 + 0.65 * h; g = 0.45 + 0.50 * h; b = 0.95 + 0.05 * h;
      }
      double du = std::fabs(u - sunU); if (du > 0.5) du = 1.0 - du;
      const double dv = v - sunV;
      const double d2 = (du * du * 4.0 + dv * dv) / (sunRadius * sunRadius);   // u spans 2 pi, v spans pi
      if (d2 < 1.0) { const double s = 60.0 * (1.0 - d2) + 20.0; r += s; g += 0.95 * s; b += 0.85 * s; }
      float* p = &rgba[4 * ((size_t)y * width + x)];
      p[0] = (float)r; p[1] = (float)g; p[2] = (float)b; p[3] = 1.0f;
    }
  }
  setTexels(width, height, rgba.data());
  return true;
}

// 3x3 Gaussian (sigma 0.5) of the RGB sum / 3, repeat in x, clamp in y (Texture.cpp:1500-1535).
static float gaussian3x3(const float* rgba, unsigned int w, unsigned int h, unsigned int x, unsigned int y)
{
  const unsigned int xs[3] = { (0 < x) ? x - 1 : w - 1, x, (x < w - 1) ? x + 1 : 0 };
  const unsigned int ys[3] = { (0 < y) ? y - 1 : y, y, (y < h - 1) ? y + 1 : y };
  auto sum = [&](unsigned int xi, unsigned int yi) { const float* p = rgba + ((size_t)w * ys[yi] + xs[xi]) * 4; return p[0] + p[1] + p[2]; };
  float intensity = sum(1, 1) * 0.619347f;
  float f = sum(1, 0);
  f += sum(0, 1);
  f += sum(2, 1);
  f += sum(1, 2);
  intensity += f * 0.0838195f;
  f = sum(0, 0);
  f += sum(2, 0);
  f += sum(0, 2);
  f += sum(2, 2);
  intensity += f * 0.0113437f;
  return intensity / 3.0f;
}

// PBRT-style piecewise-constant 2D distribution: conditional CDFs of width+1 entries per row (0 ... 1),
// marginal CDF of height+1 entries; rows weighted by sin(theta) (Texture.cpp:1540-1645).
void EnvMap::calculateSphericalCDF()
{
  const unsigned int W = m_width, H = m_height;
  std::vector<float> funcU((size_t)W * H), funcV(H + 1);
  const float* rgba = m_rgba.data();
  float sum = 0.0f;
  for (unsigned int y = 0; y < H; ++y)
  {
    const float sinTheta = float(std::sin(M_PI * (double(y) + 0.5) / double(H)));
    for (unsigned int x = 0; x < W; ++x)
    {
      funcU[(size_t)y * W + x] = gaussian3x3(rgba, W, H, x, y) * sinTheta;
      const float* p = rgba + ((size_t)y * W + x) * 4;
      sum += ((p[0] + p[1] + p[2]) / 3.0f) * sinTheta;
    }
  }
  m_integral = sum * 2.0f * RT_PI_F * RT_PI_F / float(W * H);

  m_cdfU.assign((size_t)(W + 1) * H, 0.0f);
  m_cdfV.assign(H + 1, 0.0f);
  for (unsigned int y = 0; y < H; ++y)
  {
    float* row = &m_cdfU[(size_t)y * (W + 1)];
    row[0] = 0.0f;
    for (unsigned int x = 1; x <= W; ++x) row[x] = row[x - 1] + funcU[(size_t)y * W + x - 1];
    const float integral = row[W];
    funcV[y] = integral;
    if (integral != 0.0f) { for (unsigned int x = 1; x <= W; ++x) row[x] /= integral; }
    else                  { for (unsigned int x = 1; x <= W; ++x) row[x] = float(x) / float(W); }
  }
  for (unsigned int y = 1; y <= H; ++y) m_cdfV[y] = m_cdfV[y - 1] + funcV[y - 1];
  const float integral = m_cdfV[H];
  if (integral != 0.0f) { for (unsigned int y = 1; y <= H; ++y) m_cdfV[y] /= integral; }
  else                  { for (unsigned int y = 1; y <= H; ++y) m_cdfV[y] = float(y) / float(H); }
}

// Minimal Radiance .hdr reader (-Y h +X w, RGBE, flat or adaptive RLE). Rows are flipped so row 0 is the south pole.
bool EnvMap::loadHDR(std::string const& filename)
{
  FILE* f = std::fopen(filename.c_str(), "rb");
  if (!f) return false;
  char line[512];
  bool ok = false; int w = 0, h = 0;
  while (std::fgets(line, sizeof(line), f))
  {
    if (line[0] == '\n' || line[0] == '\r') { if (std::fgets(line, sizeof(line), f) && std::sscanf(line, "-Y %d +X %d", &h, &w) == 2) ok = true; break; }
  }
  if (!ok || w <= 0 || h <= 0) { std::fclose(f); return false; }
  std::vector<float> rgba((size_t)4 * w * h);
  std::vector<unsigned char> scan((size_t)4 * w);
  for (int y = 0; y < h && ok; ++y)
  {
    unsigned char hd[4];
    if (std::fread(hd, 1, 4, f) != 4) { ok = false; break; }
    if (hd[0] == 2 && hd[1] == 2 && ((hd[2] << 8) | hd[3]) == w && w >= 8 && w < 32768)
    {
      for (int c = 0; c < 4 && ok; ++c)
        for (int x = 0; x < w && ok;)
        {
          int n = std::fgetc(f); if (n == EOF) { ok = false; break; }
          if (n > 128) { n -= 128; const int val = std::fgetc(f); if (val == EOF || x + n > w) { ok = false; break; } while (n--) scan[4 * (size_t)x++ + c] = (unsigned char)val; }
          else { if (n == 0 || x + n > w) { ok = false; break; } while (n--) { const int val = std::fgetc(f); if (val == EOF) { ok = false; break; } scan[4 * (size_t)x++ + c] = (unsigned char)val; } }
        }
    }
    else
    {
      std::memcpy(scan.data(), hd, 4);
      if (std::fread(scan.data() + 4, 1, (size_t)4 * (w - 1), f) != (size_t)4 * (w - 1)) { ok = false; break; }
    }
    float* dst = &rgba[(size_t)4 * w * (h - 1 - y)];
    for (int x = 0; x < w; ++x)
    {
      const unsigned char* p = &scan[4 * (size_t)x];
      const float s = p[3] ? std::ldexp(1.0f, (int)p[3] - 136) : 0.0f;
      dst[4 * x] = p[0] * s; dst[4 * x + 1] = p[1] * s; dst[4 * x + 2] = p[2] * s; dst[4 * x + 3] = 1.0f;
    }
  }
  std::fclose(f);
  if (!ok) return false;
  setTexels((unsigned)w, (unsigned)h, rgba.data());
  return true;
}

// ------------------------------------------------------------------------------------------------------------------
// Plain 2D pictures for the material textures
// ------------------------------------------------------------------------------------------------------------------
void EnvMap::setTexels2D(unsigned int width, unsigned int height, const float* rgba)
{
  m_width = width; m_height = height;
  m_rgba.assign(rgba, rgba + (size_t)4 * width * height);
  m_cdfU.clear(); m_cdfV.clear(); m_blob.clear();
  m_integral = 1.0f;
}

std::vector<float> const& EnvMap::getHandleBlob()
{
  if (m_blob.empty() && m_width && m_height)
  {
    m_blob.resize(4 + m_rgba.size());
    const unsigned int header[4] = { m_width, m_height, 0u, 0u };
    std::memcpy(m_blob.data(), header, sizeof(header));
    std::memcpy(m_blob.data() + 4, m_rgba.data(), m_rgba.size() * sizeof(float));
  }
  return m_blob;
}

bool EnvMap::loadImage(std::string const& filename)
{
  const size_t dot = filename.rfind('.');
  std::string ext = (dot == std::string::npos) ? std::string() : filename.substr(dot + 1);
  for (char& c : ext) c = (char)std::tolower((unsigned char)c);
  if (ext == "hdr")
  {
    if (!loadHDR(filename)) return false;
    m_cdfU.clear(); m_cdfV.clear(); m_blob.clear();
    return true;
  }
  int w = 0, h = 0; std::vector<unsigned char> bytes;
  const bool ok = (ext == "png") ? readPNG(filename, w, h, bytes) : ((ext == "ppm" || ext == "pgm" || ext == "pnm") ? readPNM(filename, w, h, bytes) : false);
  if (!ok) return false;
  std::vector<float> rgba((size_t)4 * w * h);
  for (int y = 0; y < h; ++y)      // flip: texel row 0 = bottom row of the file
    for (int x = 0; x < w; ++x)
      for (int k = 0; k < 4; ++k)
        rgba[4 * ((size_t)(h - 1 - y) * w + x) + k] = (float)bytes[4 * ((size_t)y * w + x) + k] / 255.0f;
  setTexels2D((unsigned int)w, (unsigned int)h, rgba.data());
  return true;
}

void EnvMap::createAlbedoProcedural(unsigned int width, unsigned int height)
{
  std::vector<float> rgba((size_t)4 * width * height);
  for (unsigned int y = 0; y < height; ++y)
    for (unsigned int x = 0; x < width; ++x)
    {
      const unsigned int tx = (x * 8u) / width, ty = (y * 8u) / height;
      const float g = 0.25f + 0.75f * (float)((x * 8u) % width) / (float)width;      // gradient inside a tile
      float* p = &rgba[4 * ((size_t)y * width + x)];
      if ((tx ^ ty) & 1u) { p[0] = 0.46f * g; p[1] = 0.73f * g; p[2] = 0.0f; }       // green tiles
      else                { p[0] = g; p[1] = g; p[2] = g; }                           // grey tiles
      p[3] = 1.0f;
    }
  setTexels2D(width, height, rgba.data());
}

void EnvMap::createCutoutProcedural(unsigned int width, unsigned int height)
{
  std::vector<float> rgba((size_t)4 * width * height);
  for (unsigned int y = 0; y < height; ++y)
    for (unsigned int x = 0; x < width; ++x)
    {
      const unsigned int band = ((y * 16u) / height) & 3u;        // 16 horizontal bands, period 4
      const unsigned int cell = (x * 64u) / width;                 // 64 columns
      float v = 1.0f;                                              // bands 0 and 2: opaque bars
      if (band == 1u) v = ((cell & 7u) >= 1u && (cell & 7u) <= 6u) ? 0.0f : 1.0f;    // slots
      else if (band == 3u) v = 0.5f;                               // half transparent: the stochastic test at work
      float* p = &rgba[4 * ((size_t)y * width + x)];
      p[0] = v; p[1] = v; p[2] = v; p[3] = 1.0f;
    }
  setTexels2D(width, height, rgba.data());
}
