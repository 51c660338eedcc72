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

  // Plain 2D pictures (the reference's hard-coded "albedo" and "cutout" Pictures, Application.cpp:679-690): texels only, no
  // CDFs.  Row 0 of the texel array is v = 0, i.e. the BOTTOM row of an image file (DevIL's lower-left origin, Picture.cpp).
  void setTexels2D(unsigned int width, unsigned int height, const float* rgba);
  bool loadImage(std::string const& filename);         // .png, .pgm/.ppm, .hdr
  void createAlbedoProcedural(unsigned int width, unsigned int height);   // two-tone checker with a per-tile gradient
  void createCutoutProcedural(unsigned int width, unsigned int height);   // slots: opaque bars, holes, and a half-transparent band
  // the texels behind a 16-byte header {width, height, 0, 0}: the layout of a material-texture handle (rtc_texture_create,
  // oracle/rt_oracle.c tex2d_wrap); the address of the blob is a valid HOST handle for the CPU checker
  std::vector<float> const& getHandleBlob();

  unsigned int getWidth() const { return m_width; }
  unsigned int getHeight() const { return m_height; }
  float getIntegral() const { return m_integral; }
  std::vector<float> const& getTexels() const { return m_rgba; }
  std::vector<float> const& getCDF_U() const { return m_cdfU; }
  std::vector<float> const& getCDF_V() const { return m_cdfV; }

private:
  unsigned int m_width = 0, m_height = 0;
  float m_integral = 1.0f;
  std::vector<float> m_rgba, m_cdfU, m_cdfV, m_blob;
};
