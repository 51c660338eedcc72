// Transform.h -- the transform stack arithmetic of the scene description loader.
// The reference uses nvpro-pipeline's dp::math (apps/rtigo3/dp/math/Matmnt.h:1005-1020 matrix product,
// :1099-1113 quaternion -> matrix, Quatt.h:326-333 axis/angle -> quaternion): ROW-vector convention,
// matrices multiply from the right in file order ("first written = first applied", Application.cpp:1410-1413),
// translation lives in row 3.  Only what `rotate/scale/translate/push/pop/identity` need is restated here.
#pragma once
#include <cmath>

struct Mat44
{
  float m[4][4];
  static Mat44 identity()
  {
    Mat44 r;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = (i == j) ? 1.0f : 0.0f;
    return r;
  }
  // this * rhs, accumulated in float from zero in index order like Matmnt's operator*
  Mat44 operator*(Mat44 const& rhs) const
  {
    Mat44 r;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j)
      {
        float s = 0.0f;
        for (int l = 0; l < 4; ++l) s += m[i][l] * rhs.m[l][j];
        r.m[i][j] = s;
      }
    return r;
  }
  static Mat44 scaling(float x, float y, float z) { Mat44 r = identity(); r.m[0][0] = x; r.m[1][1] = y; r.m[2][2] = z; return r; }
  static Mat44 translation(float x, float y, float z) { Mat44 r = identity(); r.m[3][0] = x; r.m[3][1] = y; r.m[3][2] = z; return r; }
  // axis is normalised by the caller; angle in radians
  static Mat44 rotation(const float axis[3], float angle)
  {
    const float s = std::sin(0.5f * angle);
    const float x = axis[0] * s, y = axis[1] * s, z = axis[2] * s, w = std::cos(0.5f * angle);
    Mat44 r = identity();
    r.m[0][0] = 1 - 2 * (y * y + z * z); r.m[0][1] = 2 * (x * y + z * w);     r.m[0][2] = 2 * (x * z - y * w);
    r.m[1][0] = 2 * (x * y - z * w);     r.m[1][1] = 1 - 2 * (x * x + z * z); r.m[1][2] = 2 * (y * z + x * w);
    r.m[2][0] = 2 * (x * z + y * w);     r.m[2][1] = 2 * (y * z - x * w);     r.m[2][2] = 1 - 2 * (x * x + y * y);
    return r;
  }
  // transposed upper 3x4: the row-major object->world matrix the instances carry (Application.cpp:1354-1359)
  void toTrafo(float out[12]) const
  {
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) out[4 * r + c] = m[c][r];
  }
};

// 3x4 * 3x4 (implicit last row 0 0 0 1), Device.cpp multiplyMatrix used by traverseNode (Device.cpp:1303)
inline void multiplyMatrix(float out[12], const float a[12], const float b[12])
{
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c)
    {
      float s = a[4 * r + 0] * b[0 + c] + a[4 * r + 1] * b[4 + c] + a[4 * r + 2] * b[8 + c];
      if (c == 3) s += a[4 * r + 3];
      out[4 * r + c] = s;
    }
}
