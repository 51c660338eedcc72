// Shapes.cpp -- procedural tessellators of the rtigo3 scene graph.  See SceneGraph.h for the contract:
// the enumeration order of vertices and triangles matches the reference's so primitive ids agree.
#include "SceneGraph.h"

#include <cmath>

namespace sg
{
  static const float kPi = RT_PI_F;

  static TriangleAttributes make_attrib(float3 vertex, float3 tangent, float3 normal, float3 texcoord)
  {
    TriangleAttributes a;
    a.vertex = vertex; a.tangent = tangent; a.normal = normal; a.texcoord = texcoord;
    return a;
  }

  // Two triangles per grid cell over a (cellsU + 1)-column vertex lattice, row by row:
  // (lower left, lower right, upper right), (upper right, upper left, lower left).
  // This is the index pattern of the plane, the sphere and the torus (Plane.cpp:120-133, Sphere.cpp:88-101, Torus.cpp:96-108).
  void Triangles::gridIndices(unsigned int cellsU, unsigned int cellsV)
  {
    const unsigned int columns = cellsU + 1;
    m_indices.reserve(m_indices.size() + 6u * (size_t)cellsU * cellsV);
    for (unsigned int row = 0; row < cellsV; ++row)
    {
      for (unsigned int col = 0; col < cellsU; ++col)
      {
        const unsigned int ll = row * columns + col, lr = ll + 1, ul = ll + columns, ur = ul + 1;
        const unsigned int cell[6] = { ll, lr, ur, ur, ul, ll };
        m_indices.insert(m_indices.end(), cell, cell + 6);
      }
    }
  }

  // Axis aligned cube [-1, 1]^3, one quad per face in the order -x, +x, -z, +z, -y, +y; each quad
  // enumerates its corners counter-clockwise seen from outside with texcoords (0,0) (1,0) (1,1) (0,1).
  void Triangles::createBox()
  {
    m_attributes.clear();
    m_indices.clear();
    struct Face { float3 normal, tangent, origin, du, dv; };   // corner(s,t) = origin + s*du + t*dv
    const Face faces[6] =
    {
      { { -1, 0, 0 }, { 0, 0,  1 }, { -1, -1, -1 }, { 0, 0,  2 }, { 0, 2, 0 } },
      { {  1, 0, 0 }, { 0, 0, -1 }, {  1, -1,  1 }, { 0, 0, -2 }, { 0, 2, 0 } },
      { { 0, 0, -1 }, { -1, 0, 0 }, {  1, -1, -1 }, { -2, 0, 0 }, { 0, 2, 0 } },
      { { 0, 0,  1 }, {  1, 0, 0 }, { -1, -1,  1 }, {  2, 0, 0 }, { 0, 2, 0 } },
      { { 0, -1, 0 }, {  1, 0, 0 }, { -1, -1, -1 }, {  2, 0, 0 }, { 0, 0,  2 } },
      { { 0,  1, 0 }, {  1, 0, 0 }, { -1,  1,  1 }, {  2, 0, 0 }, { 0, 0, -2 } },
    };
    const float st[4][2] = { { 0, 0 }, { 1, 0 }, { 1, 1 }, { 0, 1 } };
    for (unsigned int f = 0; f < 6; ++f)
    {
      for (int c = 0; c < 4; ++c)
      {
        const float3 p = faces[f].origin + faces[f].du * st[c][0] + faces[f].dv * st[c][1];
        m_attributes.push_back(make_attrib(p, faces[f].tangent, faces[f].normal, make_float3(st[c][0], st[c][1], 0.0f)));
      }
      const unsigned int b = f * 4;
      const unsigned int quad[6] = { b, b + 1, b + 2, b + 2, b + 3, b };
      m_indices.insert(m_indices.end(), quad, quad + 6);
    }
  }

  // Plane [-1, 1]^2 through the origin with normal +upAxis, (tessU + 1) x (tessV + 1) vertices, texcoord (0,0) at the
  // lower front / left front / lower left corner for upAxis 0 / 1 / 2.
  void Triangles::createPlane(unsigned int tessU, unsigned int tessV, unsigned int upAxis)
  {
    m_attributes.clear();
    m_indices.clear();
    if (tessU < 1) tessU = 1;
    if (tessV < 1) tessV = 1;
    const float uTile = 2.0f / float(tessU);
    const float vTile = 2.0f / float(tessV);
    float3 corner, tangent, normal, stepU, stepV;   // vertex = corner + u * stepU + v * stepV
    switch (upAxis)
    {
      case 0:  corner = make_float3(0.0f, -1.0f, 1.0f);  tangent = make_float3(0.0f, 0.0f, -1.0f); normal = make_float3(1.0f, 0.0f, 0.0f);
               stepU = make_float3(0.0f, 0.0f, -1.0f);   stepV = make_float3(0.0f, 1.0f, 0.0f); break;
      case 1:  corner = make_float3(-1.0f, 0.0f, 1.0f);  tangent = make_float3(1.0f, 0.0f, 0.0f);  normal = make_float3(0.0f, 1.0f, 0.0f);
               stepU = make_float3(1.0f, 0.0f, 0.0f);    stepV = make_float3(0.0f, 0.0f, -1.0f); break;
      case 2:  corner = make_float3(-1.0f, -1.0f, 0.0f); tangent = make_float3(1.0f, 0.0f, 0.0f);  normal = make_float3(0.0f, 0.0f, 1.0f);
               stepU = make_float3(1.0f, 0.0f, 0.0f);    stepV = make_float3(0.0f, 1.0f, 0.0f); break;
      default: return;   // the reference's switch has no default: no geometry
    }
    m_attributes.reserve((size_t)(tessU + 1) * (tessV + 1));
    for (unsigned int j = 0; j <= tessV; ++j)
    {
      const float v = float(j) * vTile;
      for (unsigned int i = 0; i <= tessU; ++i)
      {
        const float u = float(i) * uTile;
        // component-wise corner + (+-u or +-v or 0): one addition per component, exactly as the reference forms it
        const float3 offset = make_float3(stepU.x * u + stepV.x * v, stepU.y * u + stepV.y * v, stepU.z * u + stepV.z * v);
        m_attributes.push_back(make_attrib(corner + offset, tangent, normal, make_float3(u * 0.5f, v * 0.5f, 0.0f)));
      }
    }
    gridIndices(tessU, tessV);
  }

  // Latitude rings from the south pole (-y) upwards, tessV rings of tessU + 1 vertices (seam duplicated).
  void Triangles::createSphere(unsigned int tessU, unsigned int tessV, float radius, float maxTheta)
  {
    m_attributes.clear();
    m_indices.clear();
    if (tessU < 3) tessU = 3;
    if (tessV < 3) tessV = 3;
    m_attributes.reserve((size_t)(tessU + 1) * tessV);
    const float phiStep = 2.0f * kPi / float(tessU);
    const float thetaStep = maxTheta / float(tessV - 1);
    for (unsigned int lat = 0; lat < tessV; ++lat)
    {
      const float theta = float(lat) * thetaStep;
      const float sinTheta = sinf(theta), cosTheta = cosf(theta);
      const float texv = float(lat) / float(tessV - 1);
      for (unsigned int lon = 0; lon <= tessU; ++lon)
      {
        const float phi = float(lon) * phiStep;
        const float sinPhi = sinf(phi), cosPhi = cosf(phi);
        const float texu = float(lon) / float(tessU);
        const float3 n = make_float3(cosPhi * sinTheta, -cosTheta, -sinPhi * sinTheta);
        m_attributes.push_back(make_attrib(n * radius, make_float3(-sinPhi, 0.0f, -cosPhi), n, make_float3(texu, texv, 0.0f)));
      }
    }
    gridIndices(tessU, tessV - 1);
  }

  // Torus around the y-axis: ring radius innerRadius, tube radius outerRadius; (tessU + 1) x (tessV + 1) vertices.
  void Triangles::createTorus(unsigned int tessU, unsigned int tessV, float innerRadius, float outerRadius)
  {
    m_attributes.clear();
    m_indices.clear();
    if (tessU < 3) tessU = 3;
    if (tessV < 3) tessV = 3;
    m_attributes.reserve((size_t)(tessU + 1) * (tessV + 1));
    const float u = float(tessU), v = float(tessV);
    const float phiStep = 2.0f * kPi / u;
    const float thetaStep = 2.0f * kPi / v;
    for (unsigned int lat = 0; lat <= tessV; ++lat)
    {
      const float theta = float(lat) * thetaStep;
      const float sinTheta = sinf(theta), cosTheta = cosf(theta);
      const float radius = innerRadius + outerRadius * cosTheta;
      for (unsigned int lon = 0; lon <= tessU; ++lon)
      {
        const float phi = float(lon) * phiStep;
        const float sinPhi = sinf(phi), cosPhi = cosf(phi);
        m_attributes.push_back(make_attrib(make_float3(radius * cosPhi, outerRadius * sinTheta, radius * -sinPhi),
                                           make_float3(-sinPhi, 0.0f, -cosPhi),
                                           make_float3(cosPhi * cosTheta, sinTheta, -sinPhi * cosTheta),
                                           make_float3(float(lon) / u, float(lat) / v, 0.0f)));
      }
    }
    gridIndices(tessU, tessV);
  }

  // Footpoint + two spanning vectors; the quad used for the area light (Application.cpp:663-664).
  void Triangles::createParallelogram(float3 const& position, float3 const& vecU, float3 const& vecV, float3 const& normal)
  {
    m_attributes.clear();
    m_indices.clear();
    const float3 tangent = normalize(vecU);
    m_attributes.push_back(make_attrib(position,               tangent, normal, make_float3(0.0f, 0.0f, 0.0f)));
    m_attributes.push_back(make_attrib(position + vecU,        tangent, normal, make_float3(1.0f, 0.0f, 0.0f)));
    m_attributes.push_back(make_attrib(position + vecU + vecV, tangent, normal, make_float3(1.0f, 1.0f, 0.0f)));
    m_attributes.push_back(make_attrib(position + vecV,        tangent, normal, make_float3(0.0f, 1.0f, 0.0f)));
    const unsigned int quad[6] = { 0, 1, 2, 2, 3, 0 };   // corners run counter-clockwise, unlike the row-major lattices above
    m_indices.assign(quad, quad + 6);
  }
} // namespace sg
