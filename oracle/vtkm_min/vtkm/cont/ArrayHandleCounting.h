// oracle/vtkm_min -- see vtkm/Types.h (TEST INFRASTRUCTURE, VTK-m stand-in)
#ifndef oracle_vtkm_min_ArrayHandleCounting_h
#define oracle_vtkm_min_ArrayHandleCounting_h
#include <vtkm/cont/ArrayHandle.h>
namespace vtkm
{
namespace cont
{
template <typename T>
struct ArrayHandleCounting
{
  T Start, Step;
  Id N;
  ArrayHandleCounting(T start, T step, Id n)
    : Start(start)
    , Step(step)
    , N(n)
  {
  }
};
} // namespace cont
} // namespace vtkm
#endif
