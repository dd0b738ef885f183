// vtkm/cont/DataSet.h -- minimal stand-in (see vtkm/Types.h in this directory): explicit cell sets,
// coordinate systems and the data-set builder call the Cornell-box scene construction uses.
#ifndef b2pt_shim_vtkm_cont_DataSet_h
#define b2pt_shim_vtkm_cont_DataSet_h

#include <map>
#include <string>

#include <vtkm/cont/ArrayHandle.h>

namespace vtkm
{
struct TopologyElementTagPoint
{
};
struct TopologyElementTagCell
{
};
namespace cont
{

template <typename = void>
class CellSetExplicit
{
public:
  ArrayHandle<vtkm::UInt8> Shapes;
  ArrayHandle<vtkm::IdComponent> NumIndices;
  ArrayHandle<vtkm::Id> Connectivity;
  ArrayHandle<vtkm::Id> Offsets;
  vtkm::Id GetNumberOfCells() const { return Shapes.GetNumberOfValues(); }
  const ArrayHandle<vtkm::Id>& GetOffsetsArray(vtkm::TopologyElementTagPoint, vtkm::TopologyElementTagCell) const
  {
    return Offsets;
  }
};

// The reference passes cell sets around type-erased; the only concrete kind on this path is explicit.
class DynamicCellSet
{
  std::shared_ptr<CellSetExplicit<>> Set;

public:
  DynamicCellSet()
    : Set(std::make_shared<CellSetExplicit<>>())
  {
  }
  explicit DynamicCellSet(const CellSetExplicit<>& s)
    : Set(std::make_shared<CellSetExplicit<>>(s))
  {
  }
  template <typename CellSetType>
  const CellSetType& Cast() const
  {
    return *Set;
  }
  vtkm::Id GetNumberOfCells() const { return Set->GetNumberOfCells(); }
};

class CoordinateSystem
{
  ArrayHandle<vtkm::Vec<vtkm::Float32, 3>> Points;

public:
  CoordinateSystem() = default;
  void SetData(const ArrayHandle<vtkm::Vec<vtkm::Float32, 3>>& p) { Points = p; }
  class DataView
  {
    ArrayHandle<vtkm::Vec<vtkm::Float32, 3>> P;

  public:
    explicit DataView(const ArrayHandle<vtkm::Vec<vtkm::Float32, 3>>& p)
      : P(p)
    {
    }
    template <typename H>
    H Cast() const
    {
      return P;
    }
    vtkm::Id GetNumberOfValues() const { return P.GetNumberOfValues(); }
  };
  DataView GetData() const { return DataView(Points); }
  const ArrayHandle<vtkm::Vec<vtkm::Float32, 3>>& GetPoints() const { return Points; }
  vtkm::Bounds GetBounds() const
  {
    vtkm::Bounds b;
    auto p = Points.ReadPortal();
    for (vtkm::Id i = 0; i < p.GetNumberOfValues(); ++i)
    {
      auto v = p.Get(i);
      b.X.Min = std::fmin(b.X.Min, v[0]), b.X.Max = std::fmax(b.X.Max, v[0]);
      b.Y.Min = std::fmin(b.Y.Min, v[1]), b.Y.Max = std::fmax(b.Y.Max, v[1]);
      b.Z.Min = std::fmin(b.Z.Min, v[2]), b.Z.Max = std::fmax(b.Z.Max, v[2]);
    }
    return b;
  }
};

class Field
{
public:
  enum struct Association
  {
    ANY,
    POINTS,
    CELL_SET
  };
  Field() = default;
  Field(const std::string& name, Association, const ArrayHandle<vtkm::Float32>& data)
    : Name(name)
    , Data(data)
  {
  }
  std::string Name;
  ArrayHandle<vtkm::Float32> Data;
};

class ColorTable
{
};

class DataSet
{
  DynamicCellSet Cells;
  CoordinateSystem Coords;
  std::map<std::string, Field> Fields;

public:
  void SetCellSet(const DynamicCellSet& c) { Cells = c; }
  const DynamicCellSet& GetCellSet() const { return Cells; }
  void AddCoordinateSystem(const CoordinateSystem& c) { Coords = c; }
  const CoordinateSystem& GetCoordinateSystem() const { return Coords; }
  void AddField(const Field& f) { Fields[f.Name] = f; }
};

class DataSetBuilderExplicit
{
public:
  DataSet Create(const ArrayHandle<vtkm::Vec<vtkm::Float32, 3>>& coords, const ArrayHandle<vtkm::UInt8>& shapes,
                 const ArrayHandle<vtkm::IdComponent>& numIndices, const ArrayHandle<vtkm::Id>& connectivity,
                 const std::string& = "coords")
  {
    CellSetExplicit<> cs;
    cs.Shapes = shapes;
    cs.NumIndices = numIndices;
    cs.Connectivity = connectivity;
    cs.Offsets.Allocate(shapes.GetNumberOfValues() + 1);
    vtkm::Id off = 0;
    for (vtkm::Id c = 0; c < shapes.GetNumberOfValues(); ++c)
    {
      cs.Offsets.WritePortal().Set(c, off);
      off += numIndices.ReadPortal().Get(c);
    }
    cs.Offsets.WritePortal().Set(shapes.GetNumberOfValues(), off);
    DataSet ds;
    ds.SetCellSet(DynamicCellSet(cs));
    CoordinateSystem co;
    co.SetData(coords);
    ds.AddCoordinateSystem(co);
    return ds;
  }
};

} // namespace cont
} // namespace vtkm
#endif
