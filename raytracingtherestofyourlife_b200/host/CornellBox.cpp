#include "CornellBox.h"

#include <vector>

#include "b2pt_facade.h"

vtkm::cont::DataSet CornellBox::buildDataSet()
{
  std::vector<float> pts(3 * 89), texv(3 * 4), sphR(1);
  std::vector<int64_t> quadIds(5 * 22), sphPt(1), mq(22), tq(22), ms(1), ts(1);
  std::vector<int> mt(5), tt(5);
  b2pt_facade::Check(b2pt_scene_cornell(pts.data(), quadIds.data(), sphPt.data(), sphR.data(), mq.data(), tq.data(),
                                        ms.data(), ts.data(), mt.data(), tt.data(), texv.data()));
  std::vector<vec3> points, colours;
  for (size_t i = 0; i < 89; ++i)
    points.push_back(vec3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
  for (size_t i = 0; i < 4; ++i)
    colours.push_back(vec3(texv[3 * i], texv[3 * i + 1], texv[3 * i + 2]));
  tex = vtkm::cont::make_ArrayHandle(colours);
  matType = vtkm::cont::make_ArrayHandle(mt);
  texType = vtkm::cont::make_ArrayHandle(tt);
  auto toId = [](const std::vector<int64_t>& v) {
    std::vector<vtkm::Id> r(v.begin(), v.end());
    return vtkm::cont::make_ArrayHandle(r);
  };
  matIdx[0] = toId(mq), texIdx[0] = toId(tq), matIdx[1] = toId(ms), texIdx[1] = toId(ts);

  // cells in the reference's order: quads 0..11, the sphere's VERTEX cell (point 48), quads 12..21;
  // connectivity is the identity over the 89 points (CornellBox.cpp:401-402)
  std::vector<vtkm::UInt8> shapes;
  std::vector<vtkm::IdComponent> numIndices;
  std::vector<vtkm::Id> conn;
  std::vector<vtkm::Float32> cellOfPoint;
  for (int cell = 0; cell < 23; ++cell)
  {
    const bool vertex = (cell == 12);
    shapes.push_back(vertex ? vtkm::CELL_SHAPE_VERTEX : vtkm::CELL_SHAPE_QUAD);
    numIndices.push_back(vertex ? 1 : 4);
    for (int k = 0; k < (vertex ? 1 : 4); ++k)
    {
      conn.push_back(static_cast<vtkm::Id>(conn.size()));
      cellOfPoint.push_back(static_cast<vtkm::Float32>(cell));
    }
  }
  for (auto& v : cellOfPoint)
    v /= static_cast<vtkm::Float32>(cellOfPoint.size());
  coord.SetData(vtkm::cont::make_ArrayHandle(points));
  vtkm::cont::DataSetBuilderExplicit dsb;
  ds = dsb.Create(coord.GetPoints(), vtkm::cont::make_ArrayHandle(shapes), vtkm::cont::make_ArrayHandle(numIndices),
                  vtkm::cont::make_ArrayHandle(conn), "coords");
  field = vtkm::cont::make_ArrayHandle(cellOfPoint);
  ds.AddField(vtkm::cont::Field("point_var", vtkm::cont::Field::Association::POINTS, field));
  return ds;
}

void CornellBox::extract()
{
  const auto& cs = ds.GetCellSet().Cast<vtkm::cont::CellSetExplicit<>>();
  std::vector<vtkm::Id> spheres;
  std::vector<vtkm::Vec<vtkm::Id, 5>> quads;
  for (vtkm::Id c = 0; c < cs.GetNumberOfCells(); ++c)
  {
    const vtkm::Id off = cs.Offsets.ReadPortal().Get(c);
    auto conn = cs.Connectivity.ReadPortal();
    if (cs.Shapes.ReadPortal().Get(c) == vtkm::CELL_SHAPE_VERTEX)
      spheres.push_back(conn.Get(off));
    else
      quads.push_back(vtkm::Vec<vtkm::Id, 5>(c, conn.Get(off), conn.Get(off + 1), conn.Get(off + 2), conn.Get(off + 3)));
  }
  SphereIds = vtkm::cont::make_ArrayHandle(spheres);
  SphereRadii = vtkm::cont::make_ArrayHandle(std::vector<vtkm::Float32>(spheres.size(), static_cast<float>(90.0 / 555.0)));
  ShapeOffset = cs.Offsets;
  QuadIds = vtkm::cont::make_ArrayHandle(quads);
}
