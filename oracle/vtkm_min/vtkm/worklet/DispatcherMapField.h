// oracle/vtkm_min -- see vtkm/Types.h.  Declared only: the harness never dispatches through VTK-m.
#ifndef oracle_vtkm_min_DispatcherMapField_h
#define oracle_vtkm_min_DispatcherMapField_h
#include <vtkm/worklet/WorkletMapField.h>
namespace vtkm
{
namespace worklet
{
template <typename W>
struct DispatcherMapField
{
  DispatcherMapField() {}
  explicit DispatcherMapField(const W&) {}
  template <typename... A>
  void Invoke(A&&...) const;
};
} // namespace worklet
} // namespace vtkm
#endif
