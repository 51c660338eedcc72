// SceneGraph.h -- the host scene graph of rtigo3 (apps/rtigo3/inc/SceneGraph.h:44-142): a Group holds
// children, an Instance holds a 3x4 transform, a material id, a light id and one child, Triangles hold
// TriangleAttributes[] + uint indices.  Vertex and index ORDER of the procedural shapes defines the
// primitive ids the hit records report, so the tessellators reproduce the reference's enumeration
// (Box.cpp:37-183, Plane.cpp:37-134, Sphere.cpp:37-104, Torus.cpp:49-109, Parallelogram.cpp:46-80).
#pragma once
#include <memory>
#include <vector>

#include "HostTypes.h"

namespace sg
{
  enum NodeType { NT_GROUP, NT_INSTANCE, NT_TRIANGLES };

  class Node
  {
  public:
    explicit Node(unsigned int id) : m_id(id) {}
    virtual ~Node() {}
    virtual NodeType getType() const = 0;
    unsigned int getId() const { return m_id; }
  private:
    unsigned int m_id;
  };

  class Triangles : public Node
  {
  public:
    explicit Triangles(unsigned int id) : Node(id) {}
    NodeType getType() const override { return NT_TRIANGLES; }

    void createBox();
    void createPlane(unsigned int tessU, unsigned int tessV, unsigned int upAxis);
    void createSphere(unsigned int tessU, unsigned int tessV, float radius, float maxTheta);
    void createTorus(unsigned int tessU, unsigned int tessV, float innerRadius, float outerRadius);
    void createParallelogram(float3 const& position, float3 const& vecU, float3 const& vecV, float3 const& normal);

    void setAttributes(std::vector<TriangleAttributes> const& a) { m_attributes = a; }
    std::vector<TriangleAttributes> const& getAttributes() const { return m_attributes; }
    void setIndices(std::vector<unsigned int> const& i) { m_indices = i; }
    std::vector<unsigned int> const& getIndices() const { return m_indices; }

  private:
    void gridIndices(unsigned int cellsU, unsigned int cellsV);
    std::vector<TriangleAttributes> m_attributes;
    std::vector<unsigned int>       m_indices;
  };

  class Instance : public Node
  {
  public:
    explicit Instance(unsigned int id) : Node(id)
    {
      static const float identity[12] = { 1, 0, 0, 0,  0, 1, 0, 0,  0, 0, 1, 0 };
      setTransform(identity);
    }
    NodeType getType() const override { return NT_INSTANCE; }
    void setTransform(const float m[12]) { for (int i = 0; i < 12; ++i) m_matrix[i] = m[i]; }
    const float* getTransform() const { return m_matrix; }
    void setChild(std::shared_ptr<Node> node) { m_child = node; }
    std::shared_ptr<Node> getChild() const { return m_child; }
    void setMaterial(int index) { m_material = index; }
    int getMaterial() const { return m_material; }
    void setLight(int index) { m_light = index; }
    int getLight() const { return m_light; }
  private:
    float m_matrix[12];
    int   m_material = -1;
    int   m_light = -1;
    std::shared_ptr<Node> m_child;
  };

  class Group : public Node
  {
  public:
    explicit Group(unsigned int id) : Node(id) {}
    NodeType getType() const override { return NT_GROUP; }
    void addChild(std::shared_ptr<Instance> instance) { m_children.push_back(instance); }
    size_t getNumChildren() const { return m_children.size(); }
    std::shared_ptr<Instance> getChild(size_t index) const { return m_children[index]; }
  private:
    std::vector<std::shared_ptr<Instance>> m_children;
  };
} // namespace sg
