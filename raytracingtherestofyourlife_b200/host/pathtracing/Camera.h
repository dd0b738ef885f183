// pathtracing/Camera.h -- facade of vtkm::rendering::pathtracing::Camera (reference pathtracing/Camera.h:36-168,
// Camera.cxx).  Keeps the public surface -- parameters, validation with the reference's error messages, the
// public per-pixel RNG `seeds` array and CreateRays -- and generates rays with the sm_100a k_create_rays kernel
// through b2pt_create_rays.  PixelData / Ortho2D / subset / debug-ray code is dead on the path-tracing path
// (SURVEY.md 2): signatures are kept, GetPixelData and CreateDebugRay report that.
#ifndef b2pt_facade_pathtracing_Camera_h
#define b2pt_facade_pathtracing_Camera_h

#include <string>

#include <vtkm/rendering/Rendering.h>

#include "../raytracing/Ray.h"

namespace vtkm
{
namespace rendering
{
namespace pathtracing
{

class Camera
{
  vtkm::Int32 Height = 500, Width = 500;
  vtkm::Int32 SubsetWidth = 500, SubsetHeight = 500, SubsetMinX = 0, SubsetMinY = 0;
  vtkm::Float32 FovX = 30.f, FovY = 30.f, Zoom = 1.f;
  bool IsViewDirty = true;
  vtkm::Vec<vtkm::Float32, 3> Look{ 0.f, 0.f, -1.f }, Up{ 0.f, 1.f, 0.f }, LookAt{ 0.f, 0.f, -1.f },
    Position{ 0.f, 0.f, 0.f };
  vtkm::rendering::Camera CameraView;

public:
  // per-pixel wang-hash RNG state, read and advanced by CreateRays (reference Camera.h:62)
  vtkm::cont::ArrayHandle<unsigned int> seeds;

  Camera();
  ~Camera();

  std::string ToString();
  void SetParameters(const vtkm::rendering::Camera& camera, vtkm::rendering::CanvasRayTracer& canvas);
  void SetHeight(const vtkm::Int32& height);
  vtkm::Int32 GetHeight() const;
  void SetWidth(const vtkm::Int32& width);
  vtkm::Int32 GetWidth() const;
  vtkm::Int32 GetSubsetWidth() const;
  vtkm::Int32 GetSubsetHeight() const;
  void SetZoom(const vtkm::Float32& zoom);
  vtkm::Float32 GetZoom() const;
  void SetFieldOfView(const vtkm::Float32& degrees);
  vtkm::Float32 GetFieldOfView() const;
  void SetUp(const vtkm::Vec<vtkm::Float32, 3>& up);
  vtkm::Vec<vtkm::Float32, 3> GetUp() const;
  void SetPosition(const vtkm::Vec<vtkm::Float32, 3>& position);
  vtkm::Vec<vtkm::Float32, 3> GetPosition() const;
  void SetLookAt(const vtkm::Vec<vtkm::Float32, 3>& lookAt);
  vtkm::Vec<vtkm::Float32, 3> GetLookAt() const;
  void ResetIsViewDirty();
  bool GetIsViewDirty() const;
  void WriteSettingsToLog();

  // One jittered primary ray per pixel: directions from the per-pixel seeds (2 draws each), origin = camera
  // position, MinDistance 0, MaxDistance +inf, Distance 0, HitIdx -2, PixelIdx = j*W+i.
  void CreateRays(vtkm::rendering::raytracing::Ray<vtkm::Float32>& rays, vtkm::Bounds bounds);
  void CreateRays(vtkm::rendering::raytracing::Ray<vtkm::Float64>& rays, vtkm::Bounds bounds);
  template <typename Precision>
  void CreateRaysImpl(vtkm::rendering::raytracing::Ray<Precision>& rays, const vtkm::Bounds boundingBox);

  void GetPixelData(const vtkm::cont::CoordinateSystem& coords, vtkm::Int32& activePixels,
                    vtkm::Float32& aveRayDistance);
  void CreateDebugRay(vtkm::Vec<vtkm::Int32, 2> pixel, vtkm::rendering::raytracing::Ray<vtkm::Float32>& rays);
  void CreateDebugRay(vtkm::Vec<vtkm::Int32, 2> pixel, vtkm::rendering::raytracing::Ray<vtkm::Float64>& rays);
  bool operator==(const Camera& other) const;

private:
  void PushToDevice() const;
};

} // namespace pathtracing
} // namespace rendering
} // namespace vtkm
#endif
