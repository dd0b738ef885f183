// oracle/vtkm_min -- stub (the harness never writes data sets)
#ifndef oracle_vtkm_min_VTKDataSetWriter_h
#define oracle_vtkm_min_VTKDataSetWriter_h
#include <string>
#include <vtkm/cont/DataSet.h>
namespace vtkm
{
namespace io
{
namespace writer
{
struct VTKDataSetWriter
{
  explicit VTKDataSetWriter(const std::string&) {}
  void WriteDataSet(const vtkm::cont::DataSet&) const {}
};
}
}
}
#endif
