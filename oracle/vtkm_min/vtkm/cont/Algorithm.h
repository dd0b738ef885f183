#include <vtkm/cont/ArrayHandle.h>
