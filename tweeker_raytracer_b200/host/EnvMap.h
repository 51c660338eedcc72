// EnvMap.h -- the spherical environment light's data side: RGBA32F lat-long texels (row 0 = south pole,
// v = theta/pi), the Gaussian-filtered conditional CDFs over u per row, the marginal CDF over v and the
// integral, computed as apps/rtigo3/src/Texture.cpp:1500-1645 does (createEnv :1300-1377).
// The reference loads an .hdr through DevIL (absent here); this class creates the map procedurally
// (sky gradient + sun disc, fixed formula) or reads a Radiance RGBE .hdr file.
#pragma once
#include <string>
#include <vector>

class EnvMap
{
public:
  bool createProcedural(unsigned int width, unsigned int height);
  bool loadHDR(std::string const& filename);           // Radiance RGBE, uncompressed or new-style RLE
  void setTexels(unsigned int width, unsigned int height, const float* rgba);
  void calculateSphericalCDF();

  unsigned int getWidth() const { return m_width; }
  unsigned int getHeight() const { return m_height; }
  float getIntegral() const { return m_integral; }
  std::vector<float> const& getTexels() const { return m_rgba; }
  std::vector<float> const& getCDF_U() const { return m_cdfU; }
  std::vector<float> const& getCDF_V() const { return m_cdfV; }

private:
  unsigned int m_width = 0, m_height = 0;
  float m_integral = 1.0f;
  std::vector<float> m_rgba, m_cdfU, m_cdfV;
};
