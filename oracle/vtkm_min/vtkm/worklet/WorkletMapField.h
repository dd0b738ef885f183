// oracle/vtkm_min -- see vtkm/Types.h (TEST INFRASTRUCTURE, VTK-m stand-in).  Only the signature tag names:
// the harness calls each worklet's operator() directly, in the argument order its ExecutionSignature declares.
#ifndef oracle_vtkm_min_WorkletMapField_h
#define oracle_vtkm_min_WorkletMapField_h
#include <vtkm/Math.h>
#include <vtkm/VectorAnalysis.h>
#include <vtkm/cont/ArrayHandle.h>
namespace vtkm
{
namespace worklet
{
class WorkletMapField
{
public:
  struct FieldIn {};
  struct FieldOut {};
  struct FieldInOut {};
  struct WholeArrayIn {};
  struct WholeArrayOut {};
  struct WholeArrayInOut {};
  struct ExecObject {};
  struct WorkIndex {};
  struct _1 {}; struct _2 {}; struct _3 {}; struct _4 {}; struct _5 {}; struct _6 {}; struct _7 {}; struct _8 {};
  struct _9 {}; struct _10 {}; struct _11 {}; struct _12 {}; struct _13 {}; struct _14 {}; struct _15 {};
  struct _16 {};
};
} // namespace worklet
} // namespace vtkm
#endif
