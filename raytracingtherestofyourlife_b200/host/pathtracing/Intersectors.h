// pathtracing/Intersectors.h -- facade of pathtracing::QuadIntersector / SphereIntersector
// (reference pathtracing/QuadIntersector.{h,cxx}, SphereIntersector.{h,cxx}): the ShapeIntersector subclasses
// MapperPathTracer owns, plus the record types its public members use.  In the reference each builds its own
// VTK-m LinearBVH and launches one traversal per shape type; here SetData only records the primitive arrays --
// the single acceleration structure and the fused trace live behind b2pt_build_bvh / the bounce kernel.
#ifndef b2pt_facade_pathtracing_Intersectors_h
#define b2pt_facade_pathtracing_Intersectors_h

#include <array>

#include <vtkm/rendering/Rendering.h>

#include "../raytracing/Ray.h"

namespace vtkm
{
namespace rendering
{
namespace pathtracing
{

// N planar float arrays viewed as one record array (the reference's ArrayHandleCompositeVector of N handles).
// Field order: HitRecord U,V,T,Nx,Ny,Nz,Px,Py,Pz; ScatterRecord Ox,Oy,Oz,Dx,Dy,Dz,Ax,Ay,Az (Record.h:4-7).
template <typename T, int N>
struct RecordView
{
  std::array<vtkm::cont::ArrayHandle<T>, N> Fields;
  RecordView() = default;
  template <typename... H>
  explicit RecordView(const H&... h)
    : Fields{ { h... } }
  {
  }
  vtkm::Id GetNumberOfValues() const { return Fields[0].GetNumberOfValues(); }
  vtkm::Vec<T, N> Get(vtkm::Id i) const
  {
    vtkm::Vec<T, N> v;
    for (int k = 0; k < N; ++k)
      v[k] = Fields[k].ReadPortal().Get(i);
    return v;
  }
};

class QuadIntersector : public vtkm::rendering::raytracing::ShapeIntersector
{
public:
  using HitRecord = RecordView<vtkm::Float32, 9>;
  using HitId = RecordView<vtkm::Int32, 2>;
  using ScatterRecord = RecordView<vtkm::Float32, 9>;

  vtkm::cont::ArrayHandle<vtkm::Vec<vtkm::Id, 5>> QuadIds;
  vtkm::cont::ArrayHandle<vtkm::Id> MatIdx, TexIdx;
  vtkm::cont::CoordinateSystem Coords;

  void SetData(const vtkm::cont::CoordinateSystem& coords, vtkm::cont::ArrayHandle<vtkm::Vec<vtkm::Id, 5>> quadIds,
               vtkm::cont::ArrayHandle<vtkm::Id>& matIdx, vtkm::cont::ArrayHandle<vtkm::Id>& texIdx,
               vtkm::cont::ArrayHandle<vtkm::Int32>&, vtkm::cont::ArrayHandle<vtkm::Int32>&)
  {
    Coords = coords;
    QuadIds = quadIds;
    MatIdx = matIdx;
    TexIdx = texIdx;
  }
  vtkm::Id GetNumberOfShapes() const override { return QuadIds.GetNumberOfValues(); }
};

class SphereIntersector : public vtkm::rendering::raytracing::ShapeIntersector
{
public:
  vtkm::cont::ArrayHandle<vtkm::Id> PointIds, MatIdx, TexIdx;
  vtkm::cont::ArrayHandle<vtkm::Float32> Radii;
  vtkm::cont::CoordinateSystem Coords;

  void SetData(const vtkm::cont::CoordinateSystem& coords, vtkm::cont::ArrayHandle<vtkm::Id> pointIds,
               vtkm::cont::ArrayHandle<vtkm::Float32> radii, vtkm::cont::ArrayHandle<vtkm::Id>& matIdx,
               vtkm::cont::ArrayHandle<vtkm::Id>& texIdx, vtkm::cont::ArrayHandle<vtkm::Int32>&,
               vtkm::cont::ArrayHandle<vtkm::Int32>&)
  {
    Coords = coords;
    PointIds = pointIds;
    Radii = radii;
    MatIdx = matIdx;
    TexIdx = texIdx;
  }
  vtkm::Id GetNumberOfShapes() const override { return PointIds.GetNumberOfValues(); }
};

} // namespace pathtracing
} // namespace rendering
} // namespace vtkm
#endif
