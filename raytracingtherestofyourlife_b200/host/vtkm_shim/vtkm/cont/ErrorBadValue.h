// forwarding header of the minimal VTK-m stand-in (see vtkm/Types.h)
#include <vtkm/cont/ArrayHandle.h>
