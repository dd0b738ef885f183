// pathtracing/Camera.cxx -- see Camera.h.  Host side only: validation, parameter bookkeeping and the
// marshalling of Ray<> arrays to b2pt_create_rays.
#include "Camera.h"

#include <limits>
#include <sstream>
#include <vector>

#include "../b2pt_facade.h"

namespace vtkm
{
namespace rendering
{
namespace pathtracing
{

namespace
{
// wangXor.h:30-38; the default-constructed camera pre-hashes its seeds (reference Camera.cxx:604-616).
unsigned int WangStep(unsigned int& s)
{
  s = (s ^ 61u) ^ (s >> 16);
  s *= 9u;
  s = s ^ (s >> 4);
  s *= 0x27d4eb2du;
  s = s ^ (s >> 15);
  return s;
}
}

Camera::Camera()
{
  seeds.Allocate(static_cast<vtkm::Id>(Height) * Width);
  unsigned int* s = seeds.GetStorage();
  for (unsigned int i = 0; i < static_cast<unsigned int>(seeds.GetNumberOfValues()); ++i)
  {
    unsigned int idx = i;
    unsigned int val = WangStep(idx);
    WangStep(val);
    WangStep(val);
    WangStep(val);
    s[i] = val;
  }
}

Camera::~Camera() {}

void Camera::SetParameters(const vtkm::rendering::Camera& camera, vtkm::rendering::CanvasRayTracer& canvas)
{
  SetUp(camera.GetViewUp());
  SetLookAt(camera.GetLookAt());
  SetPosition(camera.GetPosition());
  SetZoom(camera.GetZoom());
  SetFieldOfView(camera.GetFieldOfView());
  SetHeight(static_cast<vtkm::Int32>(canvas.GetHeight()));
  SetWidth(static_cast<vtkm::Int32>(canvas.GetWidth()));
  CameraView = camera;
}

void Camera::SetHeight(const vtkm::Int32& height)
{
  if (height <= 0)
    throw vtkm::cont::ErrorBadValue("Camera height must be greater than zero.");
  if (Height != height)
  {
    Height = height;
    SetFieldOfView(FovY);
    seeds.Allocate(static_cast<vtkm::Id>(Height) * Width);
  }
}
vtkm::Int32 Camera::GetHeight() const { return Height; }

void Camera::SetWidth(const vtkm::Int32& width)
{
  if (width <= 0)
    throw vtkm::cont::ErrorBadValue("Camera width must be greater than zero.");
  if (Width != width)
  {
    Width = width;
    SetFieldOfView(FovY);
    seeds.Allocate(static_cast<vtkm::Id>(Height) * Width);
  }
}
vtkm::Int32 Camera::GetWidth() const { return Width; }
vtkm::Int32 Camera::GetSubsetWidth() const { return SubsetWidth; }
vtkm::Int32 Camera::GetSubsetHeight() const { return SubsetHeight; }

void Camera::SetZoom(const vtkm::Float32& zoom)
{
  if (zoom <= 0)
    throw vtkm::cont::ErrorBadValue("Camera zoom must be greater than zero.");
  if (Zoom != zoom)
  {
    IsViewDirty = true;
    Zoom = zoom;
  }
}
vtkm::Float32 Camera::GetZoom() const { return Zoom; }

void Camera::SetFieldOfView(const vtkm::Float32& degrees)
{
  if (degrees <= 0)
    throw vtkm::cont::ErrorBadValue("Camera feild of view must be greater than zero.");
  if (degrees > 180)
    throw vtkm::cont::ErrorBadValue("Camera feild of view must be less than 180.");
  // The horizontal field of view is derived from the aspect ratio here, but ray generation uses FovY for
  // both axes (reference Camera.cxx:936-938), so non-square canvases are stretched, not widened.
  vtkm::Float32 fovx = degrees;
  if (Width != Height)
  {
    const vtkm::Float32 vertical = std::tan(0.5f * (degrees * vtkm::Pi_180f()));
    const vtkm::Float32 aspect = vtkm::Float32(Width) / vtkm::Float32(Height);
    fovx = (2.0f * std::atan(aspect * vertical)) / vtkm::Pi_180f();
  }
  if (fovx != FovX || degrees != FovY)
    IsViewDirty = true;
  FovX = fovx;
  FovY = degrees;
  CameraView.SetFieldOfView(FovY);
}
vtkm::Float32 Camera::GetFieldOfView() const { return FovY; }

void Camera::SetUp(const vtkm::Vec<vtkm::Float32, 3>& up)
{
  if (Up != up)
  {
    Up = up;
    vtkm::Normalize(Up);
    IsViewDirty = true;
  }
}
vtkm::Vec<vtkm::Float32, 3> Camera::GetUp() const { return Up; }

void Camera::SetLookAt(const vtkm::Vec<vtkm::Float32, 3>& lookAt)
{
  if (LookAt != lookAt)
  {
    LookAt = lookAt;
    IsViewDirty = true;
  }
}
vtkm::Vec<vtkm::Float32, 3> Camera::GetLookAt() const { return LookAt; }

void Camera::SetPosition(const vtkm::Vec<vtkm::Float32, 3>& position)
{
  if (Position != position)
  {
    Position = position;
    IsViewDirty = true;
  }
}
vtkm::Vec<vtkm::Float32, 3> Camera::GetPosition() const { return Position; }

void Camera::ResetIsViewDirty() { IsViewDirty = false; }
bool Camera::GetIsViewDirty() const { return IsViewDirty; }
void Camera::WriteSettingsToLog() {}

std::string Camera::ToString()
{
  std::stringstream sstream;
  sstream << "------------------------------------------------------------\n";
  sstream << "Position : [" << Position[0] << "," << Position[1] << "," << Position[2] << "]\n";
  sstream << "LookAt   : [" << LookAt[0] << "," << LookAt[1] << "," << LookAt[2] << "]\n";
  sstream << "FOV_X    : " << FovX << "\n";
  sstream << "Up       : [" << Up[0] << "," << Up[1] << "," << Up[2] << "]\n";
  sstream << "Width    : " << Width << "\n";
  sstream << "Height   : " << Height << "\n";
  sstream << "------------------------------------------------------------\n";
  return sstream.str();
}

void Camera::PushToDevice() const
{
  const float pos[3] = { Position[0], Position[1], Position[2] };
  const float at[3] = { LookAt[0], LookAt[1], LookAt[2] };
  const float up[3] = { Up[0], Up[1], Up[2] };
  b2pt_facade::Check(b2pt_set_camera(b2pt_facade::Context(), pos, at, up, FovY, Width, Height));
}

template <typename Precision>
void Camera::CreateRaysImpl(vtkm::rendering::raytracing::Ray<Precision>& rays, const vtkm::Bounds)
{
  const vtkm::Id n = static_cast<vtkm::Id>(Width) * Height;
  // subset mode is always off on this path (reference Camera.cxx:1069): one ray per canvas pixel
  SubsetWidth = Width, SubsetHeight = Height, SubsetMinX = 0, SubsetMinY = 0;
  if (rays.NumRays != n)
    rays.Resize(static_cast<vtkm::Int32>(n));
  if (seeds.GetNumberOfValues() != n)
    seeds.Allocate(n);
  Look = LookAt - Position;
  vtkm::Normalize(Look);
  PushToDevice();

  std::vector<float> f(static_cast<size_t>(6 * n));
  float* d = f.data();
  b2pt_facade::Check(b2pt_create_rays(b2pt_facade::Context(), seeds.GetStorage(), d, d + n, d + 2 * n, d + 3 * n,
                                      d + 4 * n, d + 5 * n, reinterpret_cast<int64_t*>(rays.PixelIdx.GetStorage())));
  Precision* out[6] = { rays.DirX.GetStorage(),    rays.DirY.GetStorage(),    rays.DirZ.GetStorage(),
                        rays.OriginX.GetStorage(), rays.OriginY.GetStorage(), rays.OriginZ.GetStorage() };
  for (int k = 0; k < 6; ++k)
    for (vtkm::Id i = 0; i < n; ++i)
      out[k][i] = static_cast<Precision>(d[k * n + i]);
  const Precision inf = std::numeric_limits<Precision>::infinity();
  for (vtkm::Id i = 0; i < n; ++i)
  {
    rays.MaxDistance.GetStorage()[i] = inf;
    rays.MinDistance.GetStorage()[i] = 0;
    rays.Distance.GetStorage()[i] = 0;
    rays.HitIdx.GetStorage()[i] = -2;
  }
}

void Camera::CreateRays(vtkm::rendering::raytracing::Ray<vtkm::Float32>& rays, vtkm::Bounds bounds)
{
  CreateRaysImpl(rays, bounds);
}
void Camera::CreateRays(vtkm::rendering::raytracing::Ray<vtkm::Float64>& rays, vtkm::Bounds bounds)
{
  CreateRaysImpl(rays, bounds);
}

void Camera::GetPixelData(const vtkm::cont::CoordinateSystem&, vtkm::Int32&, vtkm::Float32&)
{
  throw vtkm::cont::ErrorBadValue("pathtracing::Camera::GetPixelData is not part of the path-tracing path");
}
void Camera::CreateDebugRay(vtkm::Vec<vtkm::Int32, 2>, vtkm::rendering::raytracing::Ray<vtkm::Float32>&)
{
  throw vtkm::cont::ErrorBadValue("pathtracing::Camera::CreateDebugRay is not part of the path-tracing path");
}
void Camera::CreateDebugRay(vtkm::Vec<vtkm::Int32, 2>, vtkm::rendering::raytracing::Ray<vtkm::Float64>&)
{
  throw vtkm::cont::ErrorBadValue("pathtracing::Camera::CreateDebugRay is not part of the path-tracing path");
}

bool Camera::operator==(const Camera& other) const
{
  return Height == other.Height && Width == other.Width && FovX == other.FovX && FovY == other.FovY &&
    Zoom == other.Zoom && Look == other.Look && LookAt == other.LookAt && Up == other.Up && Position == other.Position;
}

} // namespace pathtracing
} // namespace rendering
} // namespace vtkm
