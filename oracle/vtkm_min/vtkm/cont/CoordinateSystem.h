// oracle/vtkm_min -- see vtkm/Types.h.
#ifndef oracle_vtkm_min_CoordinateSystem_h
#define oracle_vtkm_min_CoordinateSystem_h
#include <vtkm/cont/ArrayHandle.h>
namespace vtkm
{
namespace cont
{
struct CoordinateSystem
{
  ArrayHandle<Vec<Float32, 3>> Points;
};
} // namespace cont
} // namespace vtkm
#endif
