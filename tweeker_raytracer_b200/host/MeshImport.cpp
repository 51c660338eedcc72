// MeshImport.cpp -- `model assimp <file>` of the scene description (apps/rtigo3/src/Assimp.cpp:47-319,
// Application.cpp:1837-1864) without the assimp library: a Wavefront OBJ (+ MTL) reader that produces what assimp's OBJ
// importer produces under the reference's post-processing steps (Triangulate | GenSmoothNormals | SortByPType):
//   * one mesh per (object/group, material) pair in file order; faces are fan-triangulated in file order, so primitive
//     ids follow the file; every face corner is its own vertex (the reference does not ask for JoinIdenticalVertices);
//   * missing normals are generated smooth: the sum of the (area-weighted) face normals of all corners of the mesh at
//     the same position, normalised; missing texture coordinates are (0,0,0); tangents come from
//     Application::calculateTangents (Application.cpp:2149-2228) because OBJ carries none;
//   * the scene graph is Group(root) -> Instance(identity) -> Group(object) -> Instance(identity, material) -> Triangles,
//     the shape traverseScene builds for assimp's root node with one child node per OBJ object (Assimp.cpp:200-319);
//   * the material is looked up by the usemtl name in the scene file's materials; a Kd in the .mtl replaces that
//     material's albedo (Assimp.cpp:283-288); unknown names fall back to "default";
//   * the whole model is cached by file name, so a second `model assimp` of the same file instances it (Assimp.cpp:49-53).
// Lines (l), points (p) and everything else SortByPType would have split away are skipped.
#include "Application.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

namespace {

struct ObjMesh
{
  std::string object, material;
  std::vector<TriangleAttributes> attributes;
  std::vector<unsigned int> indices;
  bool hasNormals = true, hasTexcoords = true;
};

inline void updateAABB(float3& lo, float3& hi, float3 const& v)
{
  lo.x = std::fmin(lo.x, v.x); lo.y = std::fmin(lo.y, v.y); lo.z = std::fmin(lo.z, v.z);
  hi.x = std::fmax(hi.x, v.x); hi.y = std::fmax(hi.y, v.y); hi.z = std::fmax(hi.z, v.z);
}

// "v", "v/vt", "v//vn", "v/vt/vn"; negative indices count from the end
bool parseCorner(const char* s, size_t nv, size_t nvt, size_t nvn, int& v, int& vt, int& vn)
{
  v = vt = vn = -1;
  char* end = nullptr;
  long a = std::strtol(s, &end, 10);
  if (end == s) return false;
  v = (int)(a < 0 ? (long)nv + a : a - 1);
  if (*end == '/')
  {
    const char* p = end + 1;
    if (*p != '/') { long b = std::strtol(p, &end, 10); if (end != p) vt = (int)(b < 0 ? (long)nvt + b : b - 1); }
    else end = const_cast<char*>(p);
    if (*end == '/') { const char* q = end + 1; long c = std::strtol(q, &end, 10); if (end != q) vn = (int)(c < 0 ? (long)nvn + c : c - 1); }
  }
  return 0 <= v && (size_t)v < nv && (vt < 0 || (size_t)vt < nvt) && (vn < 0 || (size_t)vn < nvn);
}

void readMaterialLibrary(std::string const& path, std::map<std::string, float3>& diffuse)
{
  std::ifstream in(path);
  if (!in) { std::cerr << "WARNING: createASSIMP() could not open material library " << path << std::endl; return; }
  std::string line, current;
  while (std::getline(in, line))
  {
    std::istringstream ls(line);
    std::string key; ls >> key;
    if (key == "newmtl") { ls >> current; }
    else if (key == "Kd" && !current.empty()) { float3 c = make_float3(0.0f); ls >> c.x >> c.y >> c.z; diffuse[current] = c; }
  }
}

} // namespace

// Application::calculateTangents (Application.cpp:2149-2228): an orthonormal basis around the existing normal, with the
// longest axis of the bounding box as reference direction.
void Application::calculateTangents(std::vector<TriangleAttributes>& attributes, std::vector<unsigned int> const& indices)
{
  if (indices.size() < 3) return;
  float3 aabbLo = attributes[indices[0]].vertex, aabbHi = attributes[indices[0]].vertex;
  for (size_t i = 0; i < indices.size(); ++i) updateAABB(aabbLo, aabbHi, attributes[indices[i]].vertex);
  const float3 extents = aabbHi - aabbLo;
  float f = extents.x;
  int maxComponent = 0;
  if (f < extents.y) { f = extents.y; maxComponent = 1; }
  if (f < extents.z) { maxComponent = 2; }
  float3 direction, bidirection;
  switch (maxComponent)
  {
    case 0: default: direction = make_float3(1.0f, 0.0f, 0.0f);  bidirection = make_float3(0.0f, 1.0f, 0.0f);  break;
    case 1:          direction = make_float3(0.0f, 1.0f, 0.0f);  bidirection = make_float3(0.0f, 0.0f, -1.0f); break;
    case 2:          direction = make_float3(0.0f, 0.0f, -1.0f); bidirection = make_float3(0.0f, 1.0f, 0.0f);  break;
  }
  for (size_t i = 0; i < attributes.size(); ++i)
  {
    float3 tangent = direction, bitangent = bidirection;
    const float3 normal = attributes[i].normal;
    if (0.001f < 1.0f - std::fabs(dot(normal, tangent)))
    {
      bitangent = normalize(cross(normal, tangent));
      tangent = normalize(cross(bitangent, normal));
    }
    else tangent = normalize(cross(bitangent, normal));
    attributes[i].tangent = tangent;
  }
}

std::shared_ptr<sg::Group> Application::createASSIMP(std::string const& filename)
{
  std::map<std::string, std::shared_ptr<sg::Group>>::const_iterator itGroup = m_mapGroups.find(filename);
  if (itGroup != m_mapGroups.end()) return itGroup->second;    // full model instancing under an Instance node

  std::shared_ptr<sg::Group> root(new sg::Group(m_idGroup++));
  m_mapGroups[filename] = root;                                  // also when loading fails: fail quicker next time

  std::ifstream in(filename);
  if (!in) { std::cerr << "createASSIMP() could not open " << filename << std::endl; return root; }
  const size_t dot = filename.rfind('.');
  std::string ext = (dot == std::string::npos) ? std::string() : filename.substr(dot + 1);
  for (char& c : ext) c = (char)std::tolower((unsigned char)c);
  if (ext != "obj") { std::cerr << "createASSIMP() " << filename << ": only Wavefront OBJ is supported by this build." << std::endl; return root; }
  const size_t slash = filename.find_last_of("/\\");
  const std::string directory = (slash == std::string::npos) ? std::string() : filename.substr(0, slash + 1);

  std::vector<float3> positions, normals, texcoords;
  std::map<std::string, float3> diffuse;
  std::vector<ObjMesh> meshes;
  std::map<std::pair<std::string, std::string>, size_t> meshIndex;
  std::string object = "default", material;
  std::string line;
  unsigned int lineNumber = 0;
  while (std::getline(in, line))
  {
    ++lineNumber;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    std::istringstream ls(line);
    std::string key; ls >> key;
    if (key == "v") { float3 p = make_float3(0.0f); ls >> p.x >> p.y >> p.z; positions.push_back(p); }
    else if (key == "vn") { float3 n = make_float3(0.0f); ls >> n.x >> n.y >> n.z; normals.push_back(n); }
    else if (key == "vt") { float3 t = make_float3(0.0f); ls >> t.x >> t.y; if (!(ls >> t.z)) t.z = 0.0f; texcoords.push_back(t); }
    else if (key == "o" || key == "g") { std::string name; std::getline(ls, name); const size_t b = name.find_first_not_of(" \t"); object = (b == std::string::npos) ? std::string("default") : name.substr(b); }
    else if (key == "usemtl") { ls >> material; }
    else if (key == "mtllib") { std::string lib; ls >> lib; if (!lib.empty()) readMaterialLibrary(directory + lib, diffuse); }
    else if (key == "f")
    {
      std::vector<int> v, vt, vn;
      std::string corner;
      bool ok = true;
      while (ls >> corner)
      {
        int a, b, c;
        if (!parseCorner(corner.c_str(), positions.size(), texcoords.size(), normals.size(), a, b, c)) { ok = false; break; }
        v.push_back(a); vt.push_back(b); vn.push_back(c);
      }
      if (!ok || v.size() < 3)
      {
        std::cerr << "WARNING: createASSIMP() " << filename << " (" << lineNumber << "): face skipped." << std::endl;
        continue;
      }
      const std::pair<std::string, std::string> id(object, material);
      std::map<std::pair<std::string, std::string>, size_t>::const_iterator it = meshIndex.find(id);
      if (it == meshIndex.end()) { meshIndex[id] = meshes.size(); meshes.push_back(ObjMesh()); meshes.back().object = object; meshes.back().material = material; it = meshIndex.find(id); }
      ObjMesh& mesh = meshes[it->second];
      for (size_t k = 1; k + 1 < v.size(); ++k)          // triangle fan, like aiProcess_Triangulate on convex polygons
      {
        const size_t corners[3] = { 0, k, k + 1 };
        for (size_t c : corners)
        {
          TriangleAttributes a;
          std::memset(&a, 0, sizeof(a));
          a.vertex = positions[(size_t)v[c]];
          a.tangent = make_float3(1.0f, 0.0f, 0.0f);
          if (0 <= vn[c]) a.normal = normals[(size_t)vn[c]]; else { a.normal = make_float3(0.0f, 0.0f, 1.0f); mesh.hasNormals = false; }
          if (0 <= vt[c]) a.texcoord = texcoords[(size_t)vt[c]]; else { a.texcoord = make_float3(0.0f); mesh.hasTexcoords = false; }
          mesh.indices.push_back((unsigned int)mesh.attributes.size());
          mesh.attributes.push_back(a);
        }
      }
    }
  }

  // one child group per OBJ object, in order of first appearance
  std::vector<std::string> objectOrder;
  std::map<std::string, std::shared_ptr<sg::Group>> objectGroups;
  static const float identity[12] = { 1, 0, 0, 0,  0, 1, 0, 0,  0, 0, 1, 0 };
  for (ObjMesh& mesh : meshes)
  {
    if (mesh.attributes.size() < 3) continue;
    if (!mesh.hasNormals)
    {
      // aiProcess_GenSmoothNormals: sum of the face normals (cross products, so weighted by area) over all corners at the same position
      struct Key { float x, y, z; bool operator<(Key const& o) const { return x != o.x ? x < o.x : (y != o.y ? y < o.y : z < o.z); } };
      std::map<Key, float3> sums;
      for (size_t i = 0; i + 2 < mesh.indices.size(); i += 3)
      {
        const float3 p0 = mesh.attributes[mesh.indices[i]].vertex, p1 = mesh.attributes[mesh.indices[i + 1]].vertex, p2 = mesh.attributes[mesh.indices[i + 2]].vertex;
        const float3 n = cross(p1 - p0, p2 - p0);
        for (int k = 0; k < 3; ++k)
        {
          const float3 p = mesh.attributes[mesh.indices[i + k]].vertex;
          const Key key = { p.x, p.y, p.z };
          std::map<Key, float3>::iterator its = sums.find(key);
          if (its == sums.end()) sums[key] = n; else its->second = its->second + n;
        }
      }
      for (TriangleAttributes& a : mesh.attributes)
      {
        const Key key = { a.vertex.x, a.vertex.y, a.vertex.z };
        const float3 n = sums[key];
        const float len = length(n);
        a.normal = (0.0f < len) ? n * (1.0f / len) : make_float3(0.0f, 0.0f, 1.0f);
      }
    }
    calculateTangents(mesh.attributes, mesh.indices);

    std::shared_ptr<sg::Triangles> geometry(new sg::Triangles(m_idGeometry++));
    geometry->setAttributes(mesh.attributes);
    geometry->setIndices(mesh.indices);
    m_geometries.push_back(geometry);

    std::shared_ptr<sg::Group>& group = objectGroups[mesh.object];
    if (!group)
    {
      group.reset(new sg::Group(m_idGroup++));
      objectOrder.push_back(mesh.object);
    }
    std::shared_ptr<sg::Instance> instance(new sg::Instance(m_idInstance++));
    instance->setTransform(identity);
    instance->setChild(geometry);
    int indexMaterial = -1;
    std::map<std::string, int>::const_iterator itm = m_mapMaterialReferences.find(mesh.material);
    if (itm != m_mapMaterialReferences.end())
    {
      indexMaterial = itm->second;
      std::map<std::string, float3>::const_iterator itd = diffuse.find(mesh.material);
      if (itd != diffuse.end()) m_materialsGUI[(size_t)indexMaterial].albedo = itd->second;      // Assimp.cpp:283-288
    }
    else
    {
      std::cerr << "WARNING: traverseScene() No material found for " << mesh.material << ". Trying default." << std::endl;
      std::map<std::string, int>::const_iterator itmd = m_mapMaterialReferences.find(std::string("default"));
      if (itmd != m_mapMaterialReferences.end()) indexMaterial = itmd->second;
      else std::cerr << "ERROR: loadSceneDescription() No default material found" << std::endl;
    }
    instance->setMaterial(indexMaterial);
    group->addChild(instance);
  }
  for (std::string const& name : objectOrder)
  {
    std::shared_ptr<sg::Instance> instance(new sg::Instance(m_idInstance++));
    instance->setTransform(identity);
    instance->setChild(objectGroups[name]);
    root->addChild(instance);
  }
  return root;
}
