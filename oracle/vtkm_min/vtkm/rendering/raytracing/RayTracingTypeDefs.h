#include <vtkm/Math.h>
#include <vtkm/cont/ArrayHandle.h>
