// oracle/vtkm_min -- see vtkm/Types.h (TEST INFRASTRUCTURE, VTK-m stand-in)
#ifndef oracle_vtkm_min_CoordinateSystem_h
#define oracle_vtkm_min_CoordinateSystem_h
#include <vtkm/cont/ArrayHandle.h>
namespace vtkm
{
namespace cont
{
struct CoordinateSystem
{
  ArrayHandle<Vec<Float32, 3>> Points;
  struct Data
  {
    ArrayHandle<Vec<Float32, 3>> H;
    template <typename T>
    T Cast() const
    {
      return H;
    }
  };
  void SetData(const ArrayHandle<Vec<Float32, 3>>& h) { Points = h; }
  Data GetData() const { return Data{ Points }; }
};
} // namespace cont
} // namespace vtkm
#endif
