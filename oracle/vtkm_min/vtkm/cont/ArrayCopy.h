// oracle/vtkm_min -- see vtkm/Types.h (TEST INFRASTRUCTURE, VTK-m stand-in)
#ifndef oracle_vtkm_min_ArrayCopy_h
#define oracle_vtkm_min_ArrayCopy_h
#include <vtkm/cont/ArrayHandleCounting.h>
namespace vtkm
{
namespace cont
{
template <typename T>
inline void ArrayCopy(const ArrayHandleCounting<T>& src, ArrayHandle<T>& dst)
{
  dst.Allocate(src.N);
  for (Id i = 0; i < src.N; ++i)
    dst.Vector()[static_cast<size_t>(i)] = static_cast<T>(src.Start + src.Step * static_cast<T>(i));
}
} // namespace cont
} // namespace vtkm
#endif
