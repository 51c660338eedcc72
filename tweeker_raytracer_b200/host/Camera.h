// Camera.h -- rtigo3's orbit camera reduced to what the renderer consumes: the five parameters of the
// system description (center, phi, theta, fov, distance), the aspect ratio of the render resolution, and
// getFrustum() -> P,U,V,W (apps/rtigo3/src/Camera.cpp:187-216).  Mouse interaction is out of scope.
#pragma once
#include "HostTypes.h"

class Camera
{
public:
  Camera();
  void setResolution(int w, int h);
  void markDirty() { m_changed = true; }
  bool getFrustum(float3& p, float3& u, float3& v, float3& w, bool force = false);
  float getAspectRatio() const { return m_aspect; }

public: // the system description loader writes these directly, as the reference does (Application.cpp:1218-1235)
  float3 m_center;
  float  m_distance;
  float  m_phi;
  float  m_theta;
  float  m_fov;

private:
  int   m_widthResolution, m_heightResolution;
  float m_aspect;
  bool  m_changed;
  float3 m_cameraP, m_cameraU, m_cameraV, m_cameraW;
};
