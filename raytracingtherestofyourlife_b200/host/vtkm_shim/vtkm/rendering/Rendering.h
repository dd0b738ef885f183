// vtkm/rendering/Rendering.h -- minimal stand-in (see vtkm/Types.h in this directory) for
// vtkm::rendering::{Camera, Canvas, CanvasRayTracer, Mapper} and raytracing::ShapeIntersector.
#ifndef b2pt_shim_vtkm_rendering_Rendering_h
#define b2pt_shim_vtkm_rendering_Rendering_h

#include <vtkm/cont/DataSet.h>

namespace vtkm
{
namespace rendering
{

class Camera
{
public:
  enum ModeEnum
  {
    MODE_2D,
    MODE_3D
  };
  using V3 = vtkm::Vec<vtkm::Float32, 3>;
  void SetPosition(const V3& p) { Position = p; }
  void SetLookAt(const V3& p) { LookAt = p; }
  void SetViewUp(const V3& p) { ViewUp = p; }
  void SetFieldOfView(vtkm::Float32 f) { FieldOfView = f; }
  void SetClippingRange(vtkm::Float32 n, vtkm::Float32 f) { Near = n, Far = f; }
  void SetZoom(vtkm::Float32 z) { Zoom = z; }
  V3 GetPosition() const { return Position; }
  V3 GetLookAt() const { return LookAt; }
  V3 GetViewUp() const { return ViewUp; }
  vtkm::Float32 GetFieldOfView() const { return FieldOfView; }
  vtkm::Float32 GetZoom() const { return Zoom; }
  ModeEnum GetMode() const { return MODE_3D; }

private:
  V3 Position{ 0.f, 0.f, 0.f }, LookAt{ 0.f, 0.f, -1.f }, ViewUp{ 0.f, 1.f, 0.f };
  vtkm::Float32 FieldOfView = 60.f, Near = 0.01f, Far = 1000.f, Zoom = 1.f;
};

class Canvas
{
public:
  Canvas(vtkm::Id w = 1024, vtkm::Id h = 1024)
    : Width(w)
    , Height(h)
  {
    Color.Allocate(w * h);
    Depth.Allocate(w * h);
  }
  virtual ~Canvas() = default;
  vtkm::Id GetWidth() const { return Width; }
  vtkm::Id GetHeight() const { return Height; }
  vtkm::cont::ArrayHandle<vtkm::Vec<vtkm::Float32, 4>>& GetColorBuffer() { return Color; }
  const vtkm::cont::ArrayHandle<vtkm::Vec<vtkm::Float32, 4>>& GetColorBuffer() const { return Color; }
  vtkm::cont::ArrayHandle<vtkm::Float32>& GetDepthBuffer() { return Depth; }

private:
  vtkm::Id Width, Height;
  vtkm::cont::ArrayHandle<vtkm::Vec<vtkm::Float32, 4>> Color;
  vtkm::cont::ArrayHandle<vtkm::Float32> Depth;
};

class CanvasRayTracer : public Canvas
{
public:
  CanvasRayTracer(vtkm::Id w = 1024, vtkm::Id h = 1024)
    : Canvas(w, h)
  {
  }
};

class Mapper
{
public:
  virtual ~Mapper() = default;
  virtual void RenderCells(const vtkm::cont::DynamicCellSet& cellset, const vtkm::cont::CoordinateSystem& coords,
                           const vtkm::cont::Field& scalarField, const vtkm::cont::ColorTable& colorTable,
                           const vtkm::rendering::Camera& camera, const vtkm::Range& scalarRange) = 0;
  virtual void SetCanvas(vtkm::rendering::Canvas* canvas) = 0;
  virtual vtkm::rendering::Canvas* GetCanvas() const = 0;
  virtual void StartScene() = 0;
  virtual void EndScene() = 0;
  virtual vtkm::rendering::Mapper* NewCopy() const = 0;
};

namespace raytracing
{
class ShapeIntersector
{
public:
  virtual ~ShapeIntersector() = default;
  virtual vtkm::Id GetNumberOfShapes() const = 0;
};
} // namespace raytracing

} // namespace rendering
} // namespace vtkm
#endif
