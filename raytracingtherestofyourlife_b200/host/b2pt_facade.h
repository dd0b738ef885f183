// b2pt_facade.h -- glue shared by the facade classes: the process-wide libb2pt contexts (one per GPU, created
// on first use) and the status -> exception mapping.  The facade calls ONLY the C-ABI of include/b2pt.h.
#ifndef b2pt_facade_h
#define b2pt_facade_h

#include <string>

#include <vtkm/cont/ArrayHandle.h>

#include "../../include/b2pt.h"

namespace b2pt_facade
{

// B2PT_ERR_BAD_VALUE is what the reference reports as vtkm::cont::ErrorBadValue; everything else (no GPU,
// CUDA failure, call order) surfaces as vtkm::cont::ErrorExecution.  There is no CPU fallback to degrade to.
inline void Check(int rc)
{
  if (rc == B2PT_OK)
    return;
  const std::string msg = b2pt_last_error();
  if (rc == B2PT_ERR_BAD_VALUE)
    throw vtkm::cont::ErrorBadValue(msg);
  throw vtkm::cont::ErrorExecution("libb2pt: " + msg);
}

// The context of one GPU; device < 0 = the default device, selected with B2PT_DEVICE (default 0).
int DefaultDevice();
b2pt_ctx* Context(int device = -1);
void ReleaseContext(); // destroys every context

} // namespace b2pt_facade
#endif
