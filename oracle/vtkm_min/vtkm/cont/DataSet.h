// oracle/vtkm_min -- see vtkm/Types.h (TEST INFRASTRUCTURE, VTK-m stand-in): just enough of DataSet / CellSetExplicit /
// Field for the reference's CornellBox.cpp to build its scene (shapes, point counts, connectivity, coordinates).
#ifndef oracle_vtkm_min_DataSet_h
#define oracle_vtkm_min_DataSet_h
#include <string>
#include <vtkm/cont/ArrayHandle.h>
#include <vtkm/cont/CoordinateSystem.h>
namespace vtkm
{
namespace cont
{
struct Field
{
  enum class Association
  {
    ANY,
    WHOLE_MESH,
    POINTS,
    CELL_SET
  };
  std::string Name;
  Field() = default;
  template <typename T>
  Field(const std::string& name, Association, const ArrayHandle<T>&)
    : Name(name)
  {
  }
};
template <typename A = void, typename B = void, typename C = void>
struct CellSetExplicit
{
  ArrayHandle<UInt8> Shapes;
  ArrayHandle<IdComponent> NumIndices;
  ArrayHandle<Id> Connectivity;
  ArrayHandle<Id> GetOffsetsArray(TopologyElementTagPoint, TopologyElementTagCell) const
  {
    ArrayHandle<Id> off;
    Id o = 0;
    for (IdComponent n : NumIndices.Vector())
    {
      off.Vector().push_back(o);
      o += n;
    }
    off.Vector().push_back(o);
    return off;
  }
};
struct DynamicCellSet
{
  CellSetExplicit<> Cells;
  template <typename T>
  T Cast() const
  {
    return Cells;
  }
};
struct DataSet
{
  DynamicCellSet CellSet;
  CoordinateSystem Coords;
  void AddField(const Field&) {}
  const DynamicCellSet& GetCellSet() const { return CellSet; }
};
} // namespace cont
} // namespace vtkm
#endif
