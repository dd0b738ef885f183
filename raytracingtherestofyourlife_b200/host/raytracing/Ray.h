// raytracing/Ray.h -- facade of vtkm::rendering::raytracing::Ray<Precision> (reference raytracing/Ray.h:48-336):
// the SoA ray container on the API of pathtracing::Camera::CreateRays and MapperPathTracer's stage methods.
// Same public members and named-buffer interface.  In the B200 design the hot loop does not round-trip
// through this container (rays live in the device-side 48-byte SoA queue); it is filled when a caller asks
// for rays explicitly (Camera::CreateRays, MapperPathTracer::intersect).
#ifndef b2pt_facade_raytracing_Ray_h
#define b2pt_facade_raytracing_Ray_h

#include <string>
#include <vector>

#include "ChannelBuffer.h"

#define RAY_ACTIVE 0
#define RAY_COMPLETE 1
#define RAY_TERMINATED 2
#define RAY_EXITED_MESH 3
#define RAY_EXITED_DOMAIN 4
#define RAY_LOST 5
#define RAY_ABANDONED 6
#define RAY_TUG_EPSILON 0.001

namespace vtkm
{
namespace rendering
{
namespace raytracing
{

// Three component arrays viewed as one Vec3 array (the reference's ArrayHandleCompositeVector members).
template <typename Precision>
struct Vec3View
{
  vtkm::cont::ArrayHandle<Precision>*X = nullptr, *Y = nullptr, *Z = nullptr;
  vtkm::Id GetNumberOfValues() const { return X ? X->GetNumberOfValues() : 0; }
  vtkm::Vec<Precision, 3> Get(vtkm::Id i) const
  {
    return vtkm::Vec<Precision, 3>(X->ReadPortal().Get(i), Y->ReadPortal().Get(i), Z->ReadPortal().Get(i));
  }
  void Set(vtkm::Id i, const vtkm::Vec<Precision, 3>& v) const
  {
    X->WritePortal().Set(i, v[0]);
    Y->WritePortal().Set(i, v[1]);
    Z->WritePortal().Set(i, v[2]);
  }
};

template <typename Precision>
class Ray
{
protected:
  bool IntersectionDataEnabled = false;

public:
  using Handle = vtkm::cont::ArrayHandle<Precision>;
  Vec3View<Precision> Intersection, Normal, Origin, Dir;
  Handle IntersectionX, IntersectionY, IntersectionZ; // hit point
  Handle OriginX, OriginY, OriginZ;
  Handle DirX, DirY, DirZ;
  Handle U, V;                      // surface parameters of the hit
  Handle NormalX, NormalY, NormalZ; // hit normal
  Handle Scalar;
  Handle Distance; // distance to hit (doubles as tmax on the path-tracing path)
  vtkm::cont::ArrayHandle<vtkm::Id> HitIdx;
  vtkm::cont::ArrayHandle<vtkm::Id> PixelIdx;
  Handle MinDistance, MaxDistance;
  vtkm::cont::ArrayHandle<vtkm::UInt8> Status; // bit 2 hit, bit 3 scattered/alive, bit 4 specular
  std::vector<ChannelBuffer<Precision>> Buffers;
  vtkm::Id DebugWidth = -1, DebugHeight = -1;
  vtkm::Id NumRays = 0;

  Ray()
  {
    Rebind();
    Buffers.emplace_back();
    Buffers.back().Resize(NumRays);
  }
  template <typename Device>
  Ray(const vtkm::Int32 size, Device, bool enableIntersectionData = false)
  {
    IntersectionDataEnabled = enableIntersectionData;
    Buffers.emplace_back();
    Resize(size);
  }
  Ray(const Ray& o) { *this = o; }
  Ray& operator=(const Ray& o)
  {
    if (this == &o)
      return *this;
    IntersectionDataEnabled = o.IntersectionDataEnabled;
    IntersectionX = o.IntersectionX, IntersectionY = o.IntersectionY, IntersectionZ = o.IntersectionZ;
    OriginX = o.OriginX, OriginY = o.OriginY, OriginZ = o.OriginZ;
    DirX = o.DirX, DirY = o.DirY, DirZ = o.DirZ;
    U = o.U, V = o.V, NormalX = o.NormalX, NormalY = o.NormalY, NormalZ = o.NormalZ;
    Scalar = o.Scalar, Distance = o.Distance, HitIdx = o.HitIdx, PixelIdx = o.PixelIdx;
    MinDistance = o.MinDistance, MaxDistance = o.MaxDistance, Status = o.Status;
    Buffers = o.Buffers;
    DebugWidth = o.DebugWidth, DebugHeight = o.DebugHeight, NumRays = o.NumRays;
    Rebind();
    return *this;
  }

  void EnableIntersectionData()
  {
    if (IntersectionDataEnabled)
      return;
    IntersectionDataEnabled = true;
    for (Handle* h : IntersectionArrays())
      h->Allocate(NumRays);
  }
  void DisableIntersectionData()
  {
    if (!IntersectionDataEnabled)
      return;
    IntersectionDataEnabled = false;
    for (Handle* h : IntersectionArrays())
      h->ReleaseResources();
  }

  void Resize(const vtkm::Int32 size)
  {
    NumRays = size;
    if (IntersectionDataEnabled)
      for (Handle* h : IntersectionArrays())
        h->Allocate(NumRays);
    for (Handle* h : { &OriginX, &OriginY, &OriginZ, &DirX, &DirY, &DirZ, &Distance, &MinDistance, &MaxDistance })
      h->Allocate(NumRays);
    Status.Allocate(NumRays);
    HitIdx.Allocate(NumRays);
    PixelIdx.Allocate(NumRays);
    Rebind();
    for (auto& b : Buffers)
      b.Resize(NumRays);
  }
  template <typename Device>
  void Resize(const vtkm::Int32 size, Device)
  {
    Resize(size);
  }

  void AddBuffer(const vtkm::Int32 numChannels, const std::string name)
  {
    ChannelBuffer<Precision> buffer(numChannels, NumRays);
    buffer.SetName(name);
    Buffers.push_back(buffer);
  }
  bool HasBuffer(const std::string name)
  {
    for (auto& b : Buffers)
      if (b.GetName() == name)
        return true;
    return false;
  }
  // The LAST buffer registered under the name wins, like the reference's lookup loop (Ray.h:293-315).
  ChannelBuffer<Precision>& GetBuffer(const std::string name)
  {
    for (size_t i = Buffers.size(); i-- > 0;)
      if (Buffers[i].GetName() == name)
        return Buffers[i];
    throw vtkm::cont::ErrorBadValue("No channel buffer with requested name: " + name);
  }

private:
  std::vector<Handle*> IntersectionArrays()
  {
    return { &IntersectionX, &IntersectionY, &IntersectionZ, &U, &V, &Scalar, &NormalX, &NormalY, &NormalZ };
  }
  void Rebind()
  {
    Intersection = { &IntersectionX, &IntersectionY, &IntersectionZ };
    Normal = { &NormalX, &NormalY, &NormalZ };
    Origin = { &OriginX, &OriginY, &OriginZ };
    Dir = { &DirX, &DirY, &DirZ };
  }
};

} // namespace raytracing
} // namespace rendering
} // namespace vtkm
#endif
