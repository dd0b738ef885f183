// forwarding header of the minimal VTK-m stand-in (see vtkm/Types.h)
#include <vtkm/rendering/Rendering.h>
