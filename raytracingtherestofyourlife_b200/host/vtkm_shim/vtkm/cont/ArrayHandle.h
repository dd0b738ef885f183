// vtkm/cont/ArrayHandle.h -- minimal stand-in (see vtkm/Types.h in this directory).
// A reference-counted, shallow-copied host array with VTK-m's portal spelling.
#ifndef b2pt_shim_vtkm_cont_ArrayHandle_h
#define b2pt_shim_vtkm_cont_ArrayHandle_h

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <vtkm/Types.h>

namespace vtkm
{
namespace cont
{

class Error : public std::runtime_error
{
public:
  explicit Error(const std::string& m)
    : std::runtime_error(m)
  {
  }
  const std::string GetMessage() const { return this->what(); }
};
class ErrorBadValue : public Error
{
public:
  explicit ErrorBadValue(const std::string& m)
    : Error(m)
  {
  }
};
class ErrorExecution : public Error
{
public:
  explicit ErrorExecution(const std::string& m)
    : Error(m)
  {
  }
};

template <typename T>
class ArrayHandle
{
  std::shared_ptr<std::vector<T>> Data;

public:
  using ValueType = T;
  class Portal
  {
    std::vector<T>* V;

  public:
    explicit Portal(std::vector<T>* v)
      : V(v)
    {
    }
    vtkm::Id GetNumberOfValues() const { return static_cast<vtkm::Id>(V->size()); }
    T Get(vtkm::Id i) const { return (*V)[static_cast<size_t>(i)]; }
    void Set(vtkm::Id i, const T& v) const { (*V)[static_cast<size_t>(i)] = v; }
  };
  ArrayHandle()
    : Data(std::make_shared<std::vector<T>>())
  {
  }
  void Allocate(vtkm::Id n) { Data->resize(static_cast<size_t>(n)); }
  void Shrink(vtkm::Id n) { Data->resize(static_cast<size_t>(n)); }
  void ReleaseResources() { Data->clear(); }
  vtkm::Id GetNumberOfValues() const { return static_cast<vtkm::Id>(Data->size()); }
  Portal ReadPortal() const { return Portal(Data.get()); }
  Portal WritePortal() const { return Portal(Data.get()); }
  // host storage access used by the facade when it hands arrays to the C-ABI
  T* GetStorage() { return Data->data(); }
  const T* GetStorage() const { return Data->data(); }
  bool SharesStorageWith(const ArrayHandle& o) const { return Data == o.Data; }
};

template <typename T>
inline ArrayHandle<T> make_ArrayHandle(const T* p, vtkm::Id n, vtkm::CopyFlag)
{
  ArrayHandle<T> h;
  h.Allocate(n);
  for (vtkm::Id i = 0; i < n; ++i)
    h.WritePortal().Set(i, p[i]);
  return h;
}
template <typename T>
inline ArrayHandle<T> make_ArrayHandle(const std::vector<T>& v, vtkm::CopyFlag f = vtkm::CopyFlag::On)
{
  return make_ArrayHandle(v.data(), static_cast<vtkm::Id>(v.size()), f);
}

} // namespace cont
} // namespace vtkm
#endif
