/* oracle/ref_scene.cxx -- TEST INFRASTRUCTURE (never linked into the product).
 *
 * Runs the reference's OWN scene builder: /root/reference/CornellBox.cpp is compiled where it lies (Makefile target
 * `ref`, second translation unit) against the VTK-m stand-in in vtkm_min/, and this file calls
 * CornellBox::buildDataSet() + CornellBox::extract() (CornellBox.cpp:141-437) and copies the arrays out in the layout
 * orc_cornell_scene uses, so tests/test_ref_harness.py can compare them bit for bit with the oracle's restatement
 * and with the library's b2pt_scene_cornell.
 *
 * What is the reference's and what is stand-in here:
 *   reference : every vertex literal, the buildQuad/buildBox point order, invert()'s matrix pipeline, material and
 *               texture index lists, the /555 normalisation, the 90/555 radius passed to the extractor.
 *   stand-in  : vtkm::Matrix / Transform3DTranslate / Transform3DRotate / MatrixMultiply (vtkm_min/vtkm/Transform3D.h,
 *               restating VTK-m's published definitions), QuadExtractor (cell id + 4 point ids per QUAD cell, in cell
 *               order) and the three SphereExtractor methods below (pathtracing/SphereExtractor.cxx needs VTK-m's
 *               topology dispatchers; its effect for an explicit cell set is: the point id of every VERTEX cell in cell
 *               order, and one constant radius per sphere).
 */
#include <cstdint>
#include <cstring>
#include <sstream>
#include <iostream>
#include "CornellBox.h"
#include "pathtracing/SphereExtractor.h"

namespace vtkm
{
namespace rendering
{
namespace pathtracing
{
void SphereExtractor::ExtractCells(const vtkm::cont::DynamicCellSet& cells, vtkm::Float32 radius)
{
  this->SetPointIdsFromCells(cells);
  this->SetUniformRadius(radius);
}
void SphereExtractor::SetPointIdsFromCells(const vtkm::cont::DynamicCellSet& cells)
{
  const auto& cs = cells.Cells;
  vtkm::Id off = 0;
  for (size_t c = 0; c < cs.Shapes.Vector().size(); ++c)
  {
    if (cs.Shapes.Vector()[c] == vtkm::CELL_SHAPE_VERTEX)
      PointIds.Vector().push_back(cs.Connectivity.Vector()[static_cast<size_t>(off)]);
    off += cs.NumIndices.Vector()[c];
  }
}
void SphereExtractor::SetUniformRadius(const vtkm::Float32 radius)
{
  Radii.Allocate(PointIds.GetNumberOfValues());
  for (auto& r : Radii.Vector())
    r = radius;
}
vtkm::cont::ArrayHandle<vtkm::Id> SphereExtractor::GetPointIds()
{
  return PointIds;
}
vtkm::cont::ArrayHandle<vtkm::Float32> SphereExtractor::GetRadii()
{
  return Radii;
}
}
}
}

/* Same signature as orc_cornell_scene (b2pt_oracle.h); additionally returns the counts through n[3] = {points, quads,
 * spheres} so a size mismatch shows up as a test failure rather than an overrun (caller passes generous buffers). */
extern "C" int b2ref_cornell_scene(float* pts, int64_t* quadIds, int64_t* sphPt, float* sphR, int64_t* matIdxQ,
                                   int64_t* texIdxQ, int64_t* matIdxS, int64_t* texIdxS, int* matType, int* texType,
                                   float* tex, int64_t* n, int64_t cap)
{
  CornellBox cb;
  cb.buildDataSet();
  std::streambuf* old = std::cout.rdbuf(); /* extract() prints the sphere ids; keep test output clean */
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf());
  cb.extract();
  std::cout.rdbuf(old);

  const auto& P = cb.ds.Coords.Points.Vector();
  const auto& Q = cb.QuadIds.Vector();
  const auto& S = cb.SphereIds.Vector();
  n[0] = static_cast<int64_t>(P.size());
  n[1] = static_cast<int64_t>(Q.size());
  n[2] = static_cast<int64_t>(S.size());
  if (n[0] > cap || n[1] > cap || n[2] > cap)
    return 1;
  if (cb.matIdx[0].Vector().size() != Q.size() || cb.texIdx[0].Vector().size() != Q.size() ||
      cb.matIdx[1].Vector().size() != S.size() || cb.texIdx[1].Vector().size() != S.size() ||
      cb.matType.Vector().size() != 5 || cb.texType.Vector().size() != 5 || cb.tex.Vector().size() != 4)
    return 2;
  for (size_t i = 0; i < P.size(); ++i)
    for (int k = 0; k < 3; ++k)
      pts[3 * i + k] = P[i][k];
  for (size_t i = 0; i < Q.size(); ++i)
  {
    for (int k = 0; k < 5; ++k)
      quadIds[5 * i + k] = Q[i][k];
    matIdxQ[i] = cb.matIdx[0].Vector()[i];
    texIdxQ[i] = cb.texIdx[0].Vector()[i];
  }
  for (size_t i = 0; i < S.size(); ++i)
  {
    sphPt[i] = S[i];
    sphR[i] = cb.SphereRadii.Vector()[i];
    matIdxS[i] = cb.matIdx[1].Vector()[i];
    texIdxS[i] = cb.texIdx[1].Vector()[i];
  }
  for (int i = 0; i < 5; ++i)
  {
    matType[i] = cb.matType.Vector()[static_cast<size_t>(i)];
    texType[i] = cb.texType.Vector()[static_cast<size_t>(i)];
  }
  for (int i = 0; i < 4; ++i)
    for (int k = 0; k < 3; ++k)
      tex[3 * i + k] = cb.tex.Vector()[static_cast<size_t>(i)][k];
  return 0;
}
