// pathtracing/PathTracer.h -- facade of vtkm::rendering::pathtracing::PathTracer (reference PathTracer.h:36-84).
// MapperPathTracer holds one but never calls Render (MapperPathTracer.cxx:82): the class is the container of
// the shape intersectors and the ray camera.  Render() in the reference is VTK-m's Phong ray-caster, which is
// not on the Monte-Carlo path and is not provided.
#ifndef b2pt_facade_pathtracing_PathTracer_h
#define b2pt_facade_pathtracing_PathTracer_h

#include <vector>

#include "Camera.h"
#include "Intersectors.h"

namespace vtkm
{
namespace rendering
{
namespace pathtracing
{

class PathTracer
{
protected:
  std::vector<vtkm::rendering::raytracing::ShapeIntersector*> Intersectors;
  Camera camera;
  vtkm::cont::Field ScalarField;
  vtkm::Id NumberOfShapes = 0;
  vtkm::cont::ArrayHandle<vtkm::Vec<vtkm::Float32, 4>> ColorMap;
  vtkm::Range ScalarRange;
  bool Shade = true;

public:
  PathTracer() = default;
  ~PathTracer() { Clear(); }
  PathTracer(const PathTracer&) = delete;
  PathTracer& operator=(const PathTracer&) = delete;

  Camera& GetCamera() { return camera; }
  // takes ownership; the intersectors are deleted by Clear() / the destructor (reference PathTracer.cxx:270-279)
  void AddShapeIntersector(vtkm::rendering::raytracing::ShapeIntersector* intersector)
  {
    NumberOfShapes += intersector->GetNumberOfShapes();
    Intersectors.push_back(intersector);
  }
  void SetField(const vtkm::cont::Field& scalarField, const vtkm::Range& scalarRange)
  {
    ScalarField = scalarField;
    ScalarRange = scalarRange;
  }
  void SetColorMap(const vtkm::cont::ArrayHandle<vtkm::Vec<vtkm::Float32, 4>>& colorMap) { ColorMap = colorMap; }
  void SetShadingOn(bool on) { Shade = on; }
  void Render(vtkm::rendering::raytracing::Ray<vtkm::Float32>&)
  {
    throw vtkm::cont::ErrorBadValue("PathTracer::Render (Phong ray casting) is not part of the path-tracing path; "
                                    "use MapperPathTracer::RenderCells");
  }
  void Render(vtkm::rendering::raytracing::Ray<vtkm::Float64>&)
  {
    throw vtkm::cont::ErrorBadValue("PathTracer::Render (Phong ray casting) is not part of the path-tracing path; "
                                    "use MapperPathTracer::RenderCells");
  }
  vtkm::Id GetNumberOfShapes() const { return NumberOfShapes; }
  void Clear()
  {
    for (auto* p : Intersectors)
      delete p;
    Intersectors.clear();
    NumberOfShapes = 0;
  }
};

} // namespace pathtracing
} // namespace rendering
} // namespace vtkm
#endif
