#include <vtkm/worklet/DispatcherMapField.h>
