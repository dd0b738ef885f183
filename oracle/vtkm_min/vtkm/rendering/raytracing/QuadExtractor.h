// oracle/vtkm_min -- see vtkm/Types.h (TEST INFRASTRUCTURE, VTK-m stand-in): QuadExtractor yields Vec<Id,5>(cell, p0..p3)
// for every quad cell in cell order.
#ifndef oracle_vtkm_min_QuadExtractor_h
#define oracle_vtkm_min_QuadExtractor_h
#include <vtkm/cont/DataSet.h>
namespace vtkm
{
namespace rendering
{
namespace raytracing
{
class QuadExtractor
{
  vtkm::cont::ArrayHandle<vtkm::Vec<vtkm::Id, 5>> QuadIds;

public:
  void ExtractCells(const vtkm::cont::DynamicCellSet& cells)
  {
    const auto& cs = cells.Cells;
    vtkm::Id off = 0;
    for (size_t c = 0; c < cs.Shapes.Vector().size(); ++c)
    {
      const vtkm::IdComponent n = cs.NumIndices.Vector()[c];
      if (cs.Shapes.Vector()[c] == vtkm::CELL_SHAPE_QUAD)
      {
        const auto& conn = cs.Connectivity.Vector();
        QuadIds.Vector().push_back(vtkm::Vec<vtkm::Id, 5>(static_cast<vtkm::Id>(c), conn[off], conn[off + 1],
                                                          conn[off + 2], conn[off + 3]));
      }
      off += n;
    }
  }
  vtkm::cont::ArrayHandle<vtkm::Vec<vtkm::Id, 5>> GetQuadIds() { return QuadIds; }
};
}
}
}
#endif
