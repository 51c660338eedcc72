#include "ImageIO.h"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
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

// ------------------------------------------------------------------------------------------------------------------
// Readers
// ------------------------------------------------------------------------------------------------------------------
static bool readFile(std::string const& path, std::vector<unsigned char>& data)
{
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  if (n <= 0) { std::fclose(f); return false; }
  data.resize((size_t)n);
  const bool ok = std::fread(data.data(), 1, (size_t)n, f) == (size_t)n;
  std::fclose(f);
  return ok;
}

static unsigned int get32(const unsigned char* p) { return ((unsigned int)p[0] << 24) | ((unsigned int)p[1] << 16) | ((unsigned int)p[2] << 8) | p[3]; }

static int paeth(int a, int b, int c)
{
  const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : ((pb <= pc) ? b : c);
}

bool readPNG(std::string const& path, int& width, int& height, std::vector<unsigned char>& rgba)
{
  std::vector<unsigned char> file;
  if (!readFile(path, file) || file.size() < 8 + 25) return false;
  const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n' };
  if (std::memcmp(file.data(), sig, 8) != 0) return false;
  unsigned int w = 0, h = 0; int depth = 0, colour = -1, interlace = 0;
  std::vector<unsigned char> idat, palette, trns;
  size_t pos = 8;
  while (pos + 12 <= file.size())
  {
    const unsigned int n = get32(&file[pos]);
    const char* type = reinterpret_cast<const char*>(&file[pos + 4]);
    if (pos + 12 + (size_t)n > file.size()) return false;
    const unsigned char* body = &file[pos + 8];
    if (std::memcmp(type, "IHDR", 4) == 0 && n >= 13) { w = get32(body); h = get32(body + 4); depth = body[8]; colour = body[9]; interlace = body[12]; }
    else if (std::memcmp(type, "PLTE", 4) == 0) palette.assign(body, body + n);
    else if (std::memcmp(type, "tRNS", 4) == 0) trns.assign(body, body + n);
    else if (std::memcmp(type, "IDAT", 4) == 0) idat.insert(idat.end(), body, body + n);
    else if (std::memcmp(type, "IEND", 4) == 0) break;
    pos += 12 + (size_t)n;
  }
  if (w == 0 || h == 0 || w > 65536 || h > 65536 || interlace != 0 || (depth != 8 && depth != 16)) return false;
  int channels = 0;
  switch (colour) { case 0: channels = 1; break; case 2: channels = 3; break; case 3: channels = 1; break; case 4: channels = 2; break; case 6: channels = 4; break; default: return false; }
  if (colour == 3 && (depth != 8 || palette.empty())) return false;
  const size_t bpp = (size_t)channels * (size_t)(depth / 8), stride = bpp * w;
  std::vector<unsigned char> raw((stride + 1) * h);
  uLongf rawSize = (uLongf)raw.size();
  if (uncompress(raw.data(), &rawSize, idat.data(), (uLong)idat.size()) != Z_OK || rawSize != raw.size()) return false;
  // undo the scanline filters in place
  std::vector<unsigned char> prev(stride, 0), line(stride);
  rgba.resize((size_t)4 * w * h);
  for (unsigned int y = 0; y < h; ++y)
  {
    const unsigned char* src = &raw[(stride + 1) * y];
    const int filter = src[0];
    for (size_t i = 0; i < stride; ++i)
    {
      const int a = (i >= bpp) ? line[i - bpp] : 0, b = prev[i], c = (i >= bpp) ? prev[i - bpp] : 0;
      int v = src[1 + i];
      switch (filter) { case 0: break; case 1: v += a; break; case 2: v += b; break; case 3: v += (a + b) / 2; break; case 4: v += paeth(a, b, c); break; default: return false; }
      line[i] = (unsigned char)v;
    }
    for (unsigned int x = 0; x < w; ++x)
    {
      const unsigned char* px = &line[bpp * x];
      const size_t step = (size_t)(depth / 8);          // 16-bit samples: keep the high byte
      unsigned char* dst = &rgba[4 * ((size_t)y * w + x)];
      switch (colour)
      {
        case 0: dst[0] = dst[1] = dst[2] = px[0]; dst[3] = 255; break;
        case 2: dst[0] = px[0]; dst[1] = px[step]; dst[2] = px[2 * step]; dst[3] = 255; break;
        case 3:
        {
          const size_t idx = px[0];
          if (3 * idx + 2 >= palette.size()) return false;
          dst[0] = palette[3 * idx]; dst[1] = palette[3 * idx + 1]; dst[2] = palette[3 * idx + 2];
          dst[3] = (idx < trns.size()) ? trns[idx] : 255;
          break;
        }
        case 4: dst[0] = dst[1] = dst[2] = px[0]; dst[3] = px[step]; break;
        case 6: dst[0] = px[0]; dst[1] = px[step]; dst[2] = px[2 * step]; dst[3] = px[3 * step]; break;
      }
    }
    prev = line;
  }
  width = (int)w; height = (int)h;
  return true;
}

// binary PGM (P5) / PPM (P6), maxval <= 255
bool readPNM(std::string const& path, int& width, int& height, std::vector<unsigned char>& rgba)
{
  std::vector<unsigned char> file;
  if (!readFile(path, file) || file.size() < 7 || file[0] != 'P' || (file[1] != '5' && file[1] != '6')) return false;
  const int channels = (file[1] == '6') ? 3 : 1;
  size_t pos = 2; int values[3] = { 0, 0, 0 };
  for (int k = 0; k < 3; ++k)
  {
    for (;;)
    {
      while (pos < file.size() && std::isspace(file[pos])) ++pos;
      if (pos < file.size() && file[pos] == '#') { while (pos < file.size() && file[pos] != '\n') ++pos; continue; }
      break;
    }
    int v = 0; bool any = false;
    while (pos < file.size() && std::isdigit(file[pos])) { v = v * 10 + (file[pos] - '0'); ++pos; any = true; }
    if (!any) return false;
    values[k] = v;
  }
  ++pos;   // the single whitespace after maxval
  const int w = values[0], h = values[1], maxval = values[2];
  if (w <= 0 || h <= 0 || maxval <= 0 || maxval > 255 || pos + (size_t)w * h * channels > file.size()) return false;
  rgba.resize((size_t)4 * w * h);
  for (size_t i = 0; i < (size_t)w * h; ++i)
  {
    const unsigned char* px = &file[pos + i * channels];
    unsigned char* dst = &rgba[4 * i];
    dst[0] = px[0]; dst[1] = px[channels == 3 ? 1 : 0]; dst[2] = px[channels == 3 ? 2 : 0]; dst[3] = 255;
  }
  width = w; height = h;
  return true;
}
