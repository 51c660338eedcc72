#include "ImageIO.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include <zlib.h>

static void put32(std::vector<unsigned char>& v, unsigned int x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }

static void chunk(std::vector<unsigned char>& out, const char* type, const unsigned char* data, size_t n)
{
  put32(out, (unsigned int)n);
  const size_t start = out.size();
  out.insert(out.end(), type, type + 4);
  if (n) out.insert(out.end(), data, data + n);
  put32(out, (unsigned int)crc32(0L, &out[start], (uInt)(n + 4)));
}

bool writePNG(std::string const& path, int width, int height, const unsigned char* rgb, bool flipY)
{
  std::vector<unsigned char> raw((size_t)height * (1 + 3 * (size_t)width));
  for (int y = 0; y < height; ++y)
  {
    const int sy = flipY ? height - 1 - y : y;
    unsigned char* row = &raw[(size_t)y * (1 + 3 * (size_t)width)];
    row[0] = 0;
    std::memcpy(row + 1, rgb + (size_t)sy * 3 * width, 3 * (size_t)width);
  }
  uLongf bound = compressBound((uLong)raw.size());
  std::vector<unsigned char> z(bound);
  if (compress2(z.data(), &bound, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
  std::vector<unsigned char> out;
  const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n' };
  out.insert(out.end(), sig, sig + 8);
  std::vector<unsigned char> ihdr;
  put32(ihdr, (unsigned int)width); put32(ihdr, (unsigned int)height);
  ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
  chunk(out, "IHDR", ihdr.data(), ihdr.size());
  chunk(out, "IDAT", z.data(), bound);
  chunk(out, "IEND", nullptr, 0);
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return false;
  const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
  std::fclose(f);
  return ok;
}

bool writeHDR(std::string const& path, int width, int height, const float* rgba, bool flipY)
{
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return false;
  std::fprintf(f, "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n", height, width);
  std::vector<unsigned char> row((size_t)4 * width);
  for (int y = 0; y < height; ++y)
  {
    const float* src = rgba + (size_t)4 * width * (flipY ? height - 1 - y : y);
    for (int x = 0; x < width; ++x)
    {
      const float r = src[4 * x], g = src[4 * x + 1], b = src[4 * x + 2];
      float m = r > g ? r : g; if (b > m) m = b;
      unsigned char* p = &row[4 * (size_t)x];
      if (!(m > 1e-32f)) { p[0] = p[1] = p[2] = p[3] = 0; continue; }
      int e; const float s = std::frexp(m, &e) * 256.0f / m;
      p[0] = (unsigned char)(r * s); p[1] = (unsigned char)(g * s); p[2] = (unsigned char)(b * s); p[3] = (unsigned char)(e + 128);
    }
    std::fwrite(row.data(), 1, row.size(), f);
  }
  std::fclose(f);
  return true;
}

bool writePFM(std::string const& path, int width, int height, const float* rgba)
{
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return false;
  std::fprintf(f, "PF\n%d %d\n-1.0\n", width, height);   // little endian, rows bottom-up like the frame itself
  std::vector<float> row((size_t)3 * width);
  for (int y = 0; y < height; ++y)
  {
    const float* src = rgba + (size_t)4 * width * y;
    for (int x = 0; x < width; ++x) { row[3 * x] = src[4 * x]; row[3 * x + 1] = src[4 * x + 1]; row[3 * x + 2] = src[4 * x + 2]; }
    std::fwrite(row.data(), sizeof(float), row.size(), f);
  }
  std::fclose(f);
  return true;
}
