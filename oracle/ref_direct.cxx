// ref_direct.cxx -- TEST INFRASTRUCTURE (never linked, imported or called by the product).
//
// Pins the oracle's restatement of the -direct G-buffer rules (oracle/b2pt_oracle.c orc_direct) to the REFERENCE'S
// OWN code, compiled from /root/reference where it lies.  The classes live inside .cxx files that need all of VTK-m
// (raytracing/RayTracerNormals.cxx, raytracing/RayTracerAlbedo.cxx, pathtracing/Camera.cxx), but each of them only
// needs value types, so the Makefile lifts exactly the class definitions
//     detail::SurfaceNormals::Shade   RayTracerNormals.cxx:47-143
//     detail::SurfaceAlbedo::Shade    RayTracerAlbedo.cxx:47-147
//     Camera::PerspectiveRayGen       pathtracing/Camera.cxx:339-423
// into generated includes under _ref/ (build output, git-ignored) and they are compiled below against the minimal
// VTK-m stand-in (oracle/vtkm_min).  What this file adds is what VTK-m's dispatcher would do: call operator() once
// with the arguments the ExecutionSignature lists.
#include <cmath>
#include <cstdint>
#include <vector>

#include <vtkm/Math.h>
#include <vtkm/VectorAnalysis.h>
#include <vtkm/cont/ArrayHandle.h>
#include <vtkm/worklet/WorkletMapField.h>

namespace refdirect
{
// minimal portals for the WholeArray arguments of the Shade worklets
struct FloatPortal
{
  float* p;
  vtkm::Id n;
  float Get(vtkm::Id i) const { return p[i]; }
  void Set(vtkm::Id i, float v) const { p[i] = v; }
  vtkm::Id GetNumberOfValues() const { return n; }
};
struct ColorMapPortal
{
  vtkm::Vec<vtkm::Float32, 4> Get(vtkm::Id) const { return vtkm::Vec<vtkm::Float32, 4>(1.f, 1.f, 1.f, 1.f); }
  vtkm::Id GetNumberOfValues() const { return 2; }
};
struct Normals
{
#include "shade_normals_extract.inc"
};
struct Albedo
{
#include "shade_albedo_extract.inc"
};
} // namespace refdirect

namespace vtkm
{
namespace rendering
{
namespace pathtracing
{
class CameraD
{
public:
  class PerspectiveRayGen;
};
#define Camera CameraD
#include "camera_perspective_extract.inc"
#undef Camera
}
}
}

extern "C" {
// which: 0 normals, 1 albedo.  lightPosition / cameraPosition / lookAt as RayTracer*::run passes them (:152-157).
void b2ref_direct_shade(int which, const float* n3, const float* p3, const float* lightPos3, const float* camPos3,
                        const float* lookAt3, float* rgba4)
{
  using V3 = vtkm::Vec<vtkm::Float32, 3>;
  const V3 n(n3[0], n3[1], n3[2]), p(p3[0], p3[1], p3[2]), lp(lightPos3[0], lightPos3[1], lightPos3[2]),
    cp(camPos3[0], camPos3[1], camPos3[2]), la(lookAt3[0], lookAt3[1], lookAt3[2]);
  refdirect::FloatPortal colors{ rgba4, 4 };
  refdirect::ColorMapPortal cmap;
  const vtkm::Id hitIdx = 0;
  const vtkm::Float32 scalar = 0.5f;
  if (which == 0)
  {
    refdirect::Normals::Shade w(lp, cp, la);
    w(hitIdx, scalar, n, p, colors, cmap, vtkm::Id(0));
  }
  else
  {
    refdirect::Albedo::Shade w(lp, cp, la);
    w(hitIdx, scalar, n, p, colors, cmap, vtkm::Id(0));
  }
}
// Camera::PerspectiveRayGen for pixel idx (fovX = fovY and zoom off as Camera::CreateRaysImpl passes them, :936-941)
void b2ref_raygen_corner(int W, int H, float fovDeg, const float* look3, const float* up3, int64_t idx, float* dir3)
{
  using V3 = vtkm::Vec<vtkm::Float32, 3>;
  vtkm::rendering::pathtracing::CameraD::PerspectiveRayGen gen(W, H, fovDeg, fovDeg, V3(look3[0], look3[1], look3[2]),
                                                              V3(up3[0], up3[1], up3[2]), 0.f, W, 0, 0);
  vtkm::Float32 dx = 0.f, dy = 0.f, dz = 0.f;
  vtkm::Id pix = 0;
  gen(vtkm::Id(idx), dx, dy, dz, pix);
  dir3[0] = dx, dir3[1] = dy, dir3[2] = dz;
}
}
