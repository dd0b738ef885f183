// oracle/vtkm_min -- see vtkm/Types.h.  Declared only.
#ifndef oracle_vtkm_min_Invoker_h
#define oracle_vtkm_min_Invoker_h
namespace vtkm
{
namespace cont
{
struct Invoker
{
  template <typename... A>
  void operator()(A&&...) const;
};
} // namespace cont
} // namespace vtkm
#endif
