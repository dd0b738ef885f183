// vtkm/Types.h -- MINIMAL STAND-IN for the VTK-m headers the facade's signatures mention.
// VTK-m is not installable in this environment (no network, not in the image); this shim provides only the
// value types, array handle, data-set and rendering-base classes that the reference's public API for the
// path-tracing path names, with VTK-m's spelling, so the facade and the reference's own driver code
// (main.cc::runPath) compile unchanged.  With a real VTK-m installation, drop this directory from the include
// path: the facade uses nothing beyond what is declared here.
#ifndef b2pt_shim_vtkm_Types_h
#define b2pt_shim_vtkm_Types_h

#include <cmath>
#include <cstdint>
#include <initializer_list>
#include <limits>

#define VTKM_CONT
#define VTKM_EXEC
#define VTKM_EXEC_CONT
#define VTKM_RENDERING_EXPORT
#define vtkmNotUsed(x)

namespace vtkm
{
using Id = long long;
using IdComponent = int;
using Int8 = signed char;
using UInt8 = unsigned char;
using Int32 = std::int32_t;
using UInt32 = std::uint32_t;
using Int64 = std::int64_t;
using Float32 = float;
using Float64 = double;

enum class CopyFlag
{
  Off = 0,
  On = 1
};

enum CellShapeIdEnum
{
  CELL_SHAPE_EMPTY = 0,
  CELL_SHAPE_VERTEX = 1,
  CELL_SHAPE_LINE = 3,
  CELL_SHAPE_TRIANGLE = 5,
  CELL_SHAPE_QUAD = 9
};

template <typename T, IdComponent N>
class Vec
{
public:
  using ComponentType = T;
  static constexpr IdComponent NUM_COMPONENTS = N;
  Vec() = default;
  explicit Vec(const T& fill)
  {
    for (IdComponent i = 0; i < N; ++i)
      c[i] = fill;
  }
  template <typename... Ts, typename = typename std::enable_if<(sizeof...(Ts) == N) && (N > 1)>::type>
  Vec(Ts... vs)
    : c{ static_cast<T>(vs)... }
  {
  }
  template <typename U>
  explicit Vec(const Vec<U, N>& o)
  {
    for (IdComponent i = 0; i < N; ++i)
      c[i] = static_cast<T>(o[i]);
  }
  Vec& operator=(const T& fill)
  {
    for (IdComponent i = 0; i < N; ++i)
      c[i] = fill;
    return *this;
  }
  T& operator[](IdComponent i) { return c[i]; }
  const T& operator[](IdComponent i) const { return c[i]; }
  bool operator==(const Vec& o) const
  {
    for (IdComponent i = 0; i < N; ++i)
      if (!(c[i] == o.c[i]))
        return false;
    return true;
  }
  bool operator!=(const Vec& o) const { return !(*this == o); }
  Vec operator+(const Vec& o) const
  {
    Vec r;
    for (IdComponent i = 0; i < N; ++i)
      r.c[i] = c[i] + o.c[i];
    return r;
  }
  Vec operator-(const Vec& o) const
  {
    Vec r;
    for (IdComponent i = 0; i < N; ++i)
      r.c[i] = c[i] - o.c[i];
    return r;
  }
  Vec operator*(const T& s) const
  {
    Vec r;
    for (IdComponent i = 0; i < N; ++i)
      r.c[i] = c[i] * s;
    return r;
  }
  // VTK-m divides Vec<T> by a Float64 scalar in double precision and narrows (used as pts/555.0)
  Vec operator/(Float64 s) const
  {
    Vec r;
    for (IdComponent i = 0; i < N; ++i)
      r.c[i] = static_cast<T>(static_cast<Float64>(c[i]) / s);
    return r;
  }

private:
  T c[N] = {};
};

using Vec3f_32 = Vec<Float32, 3>;
using Vec4f_32 = Vec<Float32, 4>;

template <typename T>
inline T Dot(const Vec<T, 3>& a, const Vec<T, 3>& b)
{
  return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}
template <typename T>
inline Vec<T, 3> Cross(const Vec<T, 3>& a, const Vec<T, 3>& b)
{
  return Vec<T, 3>(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
template <typename T>
inline void Normalize(Vec<T, 3>& v)
{
  const T r = T(1) / std::sqrt(Dot(v, v));
  v = v * r;
}
inline Float64 Pi() { return 3.14159265358979323846; }
inline Float32 Pi_180f() { return 0.01745329251994329547f; }

struct Range
{
  Float64 Min = std::numeric_limits<Float64>::infinity();
  Float64 Max = -std::numeric_limits<Float64>::infinity();
  Range() = default;
  Range(Float64 lo, Float64 hi)
    : Min(lo)
    , Max(hi)
  {
  }
};
struct Bounds
{
  Range X, Y, Z;
  Bounds() = default;
  Bounds(const Range& x, const Range& y, const Range& z)
    : X(x)
    , Y(y)
    , Z(z)
  {
  }
};
} // namespace vtkm
#endif
