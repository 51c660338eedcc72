#include "Camera.h"

#include <cmath>

Camera::Camera()
: m_center(make_float3(0.0f)), m_distance(10.0f), m_phi(0.75f), m_theta(0.6f), m_fov(60.0f)
, m_widthResolution(1), m_heightResolution(1), m_aspect(1.0f), m_changed(false)
, m_cameraP(make_float3(0.0f, 0.0f, 1.0f)), m_cameraU(make_float3(1.0f, 0.0f, 0.0f))
, m_cameraV(make_float3(0.0f, 1.0f, 0.0f)), m_cameraW(make_float3(0.0f, 0.0f, -1.0f))
{
}

void Camera::setResolution(int w, int h)
{
  if (m_widthResolution != w || m_heightResolution != h)
  {
    m_widthResolution  = (0 < w) ? w : 1;
    m_heightResolution = (0 < h) ? h : 1;
    m_aspect = float(m_widthResolution) / float(m_heightResolution);
    m_changed = true;
  }
}

// phi in [0,1] is the longitude (0.75 = +z), theta in [0,1] the polar angle from the south pole (0.5 = equator).
bool Camera::getFrustum(float3& p, float3& u, float3& v, float3& w, bool force)
{
  const bool changed = force || m_changed;
  if (changed)
  {
    const float cosPhi = cosf(m_phi * 2.0f * RT_PI_F), sinPhi = sinf(m_phi * 2.0f * RT_PI_F);
    const float cosTheta = cosf(m_theta * RT_PI_F), sinTheta = sinf(m_theta * RT_PI_F);
    const float3 outward = make_float3(cosPhi * sinTheta, -cosTheta, -sinPhi * sinTheta);
    const float tanFovHalf = tanf((m_fov * 0.5f) * RT_PI_F / 180.0f);
    m_cameraP = m_center + m_distance * outward;
    m_cameraU = m_aspect * make_float3(-sinPhi, 0.0f, -cosPhi) * tanFovHalf;
    m_cameraV = make_float3(cosTheta * cosPhi, sinTheta, cosTheta * -sinPhi) * tanFovHalf;
    m_cameraW = -outward;
    p = m_cameraP; u = m_cameraU; v = m_cameraV; w = m_cameraW;
    m_changed = false;
  }
  return changed;
}
