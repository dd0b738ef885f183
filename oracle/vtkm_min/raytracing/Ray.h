// oracle/vtkm_min -- shadows the reference's raytracing/Ray.h for BVHTraverser.h's IntersectRays declaration
// (never instantiated by the harness): storage only, no arithmetic.
#ifndef oracle_vtkm_min_Ray_h
#define oracle_vtkm_min_Ray_h
#include <vtkm/cont/ArrayHandle.h>
namespace vtkm
{
namespace rendering
{
namespace raytracing
{
template <typename Precision>
struct Ray
{
  vtkm::cont::ArrayHandle<vtkm::Vec<Precision, 3>> Origin, Dir;
  vtkm::cont::ArrayHandle<Precision> Distance, MinDistance;
  vtkm::cont::ArrayHandle<vtkm::UInt8> Status;
};
} // namespace raytracing
} // namespace rendering
} // namespace vtkm
#endif
