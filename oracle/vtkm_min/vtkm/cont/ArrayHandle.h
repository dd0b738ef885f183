// oracle/vtkm_min -- see vtkm/Types.h (TEST INFRASTRUCTURE, VTK-m stand-in): a host-only array handle with the
// portal / token spelling the reference's leaf intersectors use (Surface.h:12-28, 289-309).
#ifndef oracle_vtkm_min_ArrayHandle_h
#define oracle_vtkm_min_ArrayHandle_h
#include <memory>
#include <vector>
#include <vtkm/Types.h>
namespace vtkm
{
namespace cont
{
struct DeviceAdapterTagSerial {};
struct Token {};
struct ExecutionObjectBase {};

template <typename T>
struct ArrayPortal
{
  T* p = nullptr;
  Id n = 0;
  Id GetNumberOfValues() const { return n; }
  T Get(Id i) const { return p[i]; }
  void Set(Id i, const T& v) const { p[i] = v; }
};

template <typename T>
class ArrayHandle
{
public:
  using ValueType = T;
  template <typename Device>
  struct ExecutionTypes
  {
    using Portal = ArrayPortal<T>;
    using PortalConst = ArrayPortal<T>;
  };
  ArrayHandle()
    : data(std::make_shared<std::vector<T>>())
  {
  }
  void Allocate(Id n) { data->resize(static_cast<size_t>(n)); }
  Id GetNumberOfValues() const { return static_cast<Id>(data->size()); }
  ArrayPortal<T> Portal() const { return ArrayPortal<T>{ data->data(), static_cast<Id>(data->size()) }; }
  template <typename Device>
  ArrayPortal<T> PrepareForInput(Device, Token&) const { return Portal(); }
  template <typename Device>
  ArrayPortal<T> PrepareForInPlace(Device, Token&) const { return Portal(); }
  ArrayPortal<T> WritePortal() const { return Portal(); }
  ArrayPortal<T> ReadPortal() const { return Portal(); }
  std::vector<T>& Vector() { return *data; }
  const std::vector<T>& Vector() const { return *data; }
private:
  std::shared_ptr<std::vector<T>> data;
};
template <typename T>
inline ArrayHandle<T> make_ArrayHandle(const std::vector<T>& v, vtkm::CopyFlag)
{
  ArrayHandle<T> h;
  h.Vector() = v;
  return h;
}
} // namespace cont
} // namespace vtkm
#endif
