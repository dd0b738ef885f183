// oracle/vtkm_min -- see vtkm/Types.h.  The flat layout consumed by BVHTraverser.h:45-69, 182-221: an inner node
// is 4 x Vec4f = {xmin0 ymin0 zmin0 xmax0}{ymax0 zmax0 xmin1 ymin1}{zmin1 xmax1 ymax1 zmax1}{left right - -},
// child ids bit-cast into floats, inner child = first Vec4f index, leaf child = -(offset into Leafs) - 1,
// Leafs = [count, prim...] per leaf.  The tree itself is built by the harness (ref_harness.cxx).
#ifndef oracle_vtkm_min_BVH_h
#define oracle_vtkm_min_BVH_h
#include <vtkm/cont/ArrayHandle.h>
#include <vtkm/cont/CoordinateSystem.h>
#include <vtkm/worklet/DispatcherMapField.h> // reaches BVHTraverser.h transitively in VTK-m
namespace vtkm
{
namespace rendering
{
namespace raytracing
{
struct AABBs
{
  vtkm::cont::ArrayHandle<vtkm::Float32> xmins, ymins, zmins, xmaxs, ymaxs, zmaxs;
};
struct LinearBVH
{
  vtkm::cont::ArrayHandle<vtkm::Vec<vtkm::Float32, 4>> FlatBVH;
  vtkm::cont::ArrayHandle<vtkm::Id> Leafs;
};
} // namespace raytracing
} // namespace rendering
} // namespace vtkm
#endif
