// oracle/vtkm_min -- see vtkm/Types.h (TEST INFRASTRUCTURE, VTK-m stand-in)
#ifndef oracle_vtkm_min_DataSetBuilderExplicit_h
#define oracle_vtkm_min_DataSetBuilderExplicit_h
#include <vtkm/cont/DataSet.h>
namespace vtkm
{
namespace cont
{
struct DataSetBuilderExplicit
{
  DataSet Create(const ArrayHandle<Vec<Float32, 3>>& coords, const ArrayHandle<UInt8>& shapes,
                 const ArrayHandle<IdComponent>& numIndices, const ArrayHandle<Id>& connectivity,
                 const std::string& = "coords") const
  {
    DataSet ds;
    ds.Coords.SetData(coords);
    ds.CellSet.Cells.Shapes = shapes;
    ds.CellSet.Cells.NumIndices = numIndices;
    ds.CellSet.Cells.Connectivity = connectivity;
    return ds;
  }
};
} // namespace cont
} // namespace vtkm
#endif
