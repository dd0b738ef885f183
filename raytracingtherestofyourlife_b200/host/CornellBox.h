// CornellBox.h -- facade of the reference's CornellBox scene class (CornellBox.h / CornellBox.cpp:141-418):
// same public data members and buildDataSet()/extract(), filled from the library's host-side scene builder
// b2pt_scene_cornell (the scene is input data, not a kernel).
#ifndef b2pt_facade_CornellBox_h
#define b2pt_facade_CornellBox_h

#include <vtkm/cont/DataSet.h>

using vec3 = vtkm::Vec<vtkm::Float32, 3>;

class CornellBox
{
public:
  vtkm::cont::ArrayHandle<vec3> tex;
  vtkm::cont::ArrayHandle<vtkm::Id> matIdx[2]; // [0] quads, [1] spheres
  vtkm::cont::ArrayHandle<vtkm::Id> texIdx[2];
  vtkm::cont::ArrayHandle<int> matType, texType;
  vtkm::cont::CoordinateSystem coord;
  vtkm::cont::ArrayHandle<vtkm::Float32> field;
  vtkm::cont::ArrayHandle<vtkm::Vec<vtkm::Id, 5>> QuadIds;
  vtkm::cont::ArrayHandle<vtkm::Id> SphereIds;
  vtkm::cont::ArrayHandle<vtkm::Float32> SphereRadii;
  vtkm::cont::ArrayHandle<vtkm::Id> ShapeOffset;
  vtkm::cont::DataSet ds;

  // 89 points, 22 quad cells and 1 vertex cell (the dielectric sphere's centre), material / texture tables.
  vtkm::cont::DataSet buildDataSet();
  // fills QuadIds / SphereIds / SphereRadii / ShapeOffset from ds (CornellBox.cpp:420-437)
  void extract();
};
#endif
