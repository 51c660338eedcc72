// ImageIO.h -- writers for the screenshot path (the reference used DevIL: PNG for tonemapped, .hdr for linear output,
// Application.cpp:2253-2317).  PNG through zlib, Radiance RGBE .hdr, and PFM for lossless float dumps.
#pragma once
#include <string>

bool writePNG(std::string const& path, int width, int height, const unsigned char* rgb, bool flipY);
bool writeHDR(std::string const& path, int width, int height, const float* rgba, bool flipY);
bool writePFM(std::string const& path, int width, int height, const float* rgba);
