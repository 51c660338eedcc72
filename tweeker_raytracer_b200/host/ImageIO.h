// ImageIO.h -- writers for the screenshot path (the reference used DevIL: PNG for tonemapped, .hdr for linear output,
// Application.cpp:2253-2317).  PNG through zlib, Radiance RGBE .hdr, and PFM for lossless float dumps.
// Readers for the material pictures (the reference loaded them through DevIL, Picture.cpp): PNG (8/16-bit grey, grey+alpha,
// RGB, RGBA, palette; non-interlaced) through zlib, and binary PGM/PPM.  JPEG is not supported.
#pragma once
#include <string>
#include <vector>

bool writePNG(std::string const& path, int width, int height, const unsigned char* rgb, bool flipY);
bool writeHDR(std::string const& path, int width, int height, const float* rgba, bool flipY);
bool writePFM(std::string const& path, int width, int height, const float* rgba);

// Decodes into 8-bit RGBA, row 0 = top row of the file.  false when the file is missing, damaged or of an unsupported kind.
bool readPNG(std::string const& path, int& width, int& height, std::vector<unsigned char>& rgba);
bool readPNM(std::string const& path, int& width, int& height, std::vector<unsigned char>& rgba);
