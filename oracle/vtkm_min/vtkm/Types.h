// oracle/vtkm_min -- TEST INFRASTRUCTURE.  A minimal stand-in for the few VTK-m headers that the reference's
// header-only worklets include, so that those worklets (the reference's own arithmetic) can be compiled from
// /root/reference as they lie and driven by oracle/ref_harness.cxx.  VTK-m itself (unpinned, ~1.5.x/1.6-dev) is
// absent from this environment; the semantics restated here are the documented ones of that release line:
//   * Vec arithmetic is component-wise in T; Vec<T> (*,/) Float64 scalar computes in double and narrows
//     (vtkm/Types.h: "operator*(Vec<T,Size>, vtkm::Float64)");
//   * Dot(Vec3) accumulates left to right; Cross is the plain 6-product form (pre-1.6);
//   * Normalize(v) = v * RSqrt(MagnitudeSquared(v)), RSqrt on the host = 1/sqrt (vtkm/VectorAnalysis.h, Math.h);
//   * Min/Max on floating point = fmin/fmax (vtkm/Math.h); Epsilon<Float32>() = 1e-5f; Pi() is Float64.
#ifndef oracle_vtkm_min_Types_h
#define oracle_vtkm_min_Types_h

#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>

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
struct TopologyElementTagPoint {};
struct TopologyElementTagCell {};

template <typename T, IdComponent N>
class Vec
{
public:
  using ComponentType = T;
  static constexpr IdComponent NUM_COMPONENTS = N;
  Vec() = default;
  template <typename U, typename = typename std::enable_if<std::is_arithmetic<U>::value>::type>
  explicit Vec(const U& fill)
  {
    for (IdComponent i = 0; i < N; ++i)
      c[i] = static_cast<T>(fill);
  }
  template <typename A, typename B, typename... Ts,
            typename = typename std::enable_if<(sizeof...(Ts) + 2 == N)>::type>
  Vec(const A& a, const B& b, const Ts&... vs)
    : c{ static_cast<T>(a), static_cast<T>(b), static_cast<T>(vs)... }
  {
  }
  template <typename U>
  explicit Vec(const Vec<U, N>& o)
  {
    for (IdComponent i = 0; i < N; ++i)
      c[i] = static_cast<T>(o[i]);
  }
  T& operator[](Id i) { return c[i]; }
  const T& operator[](Id i) const { return c[i]; }
  Vec operator-() const
  {
    Vec r;
    for (IdComponent i = 0; i < N; ++i)
      r.c[i] = -c[i];
    return r;
  }
  T c[N];
};

#define ORACLE_VTKM_VEC_OP(OP)                                                                                        \
  template <typename T, IdComponent N>                                                                                \
  inline Vec<T, N> operator OP(const Vec<T, N>& a, const Vec<T, N>& b)                                                \
  {                                                                                                                   \
    Vec<T, N> r;                                                                                                      \
    for (IdComponent i = 0; i < N; ++i)                                                                               \
      r[i] = a[i] OP b[i];                                                                                            \
    return r;                                                                                                         \
  }                                                                                                                   \
  template <typename T, IdComponent N>                                                                                \
  inline Vec<T, N> operator OP(const Vec<T, N>& a, T s)                                                               \
  {                                                                                                                   \
    Vec<T, N> r;                                                                                                      \
    for (IdComponent i = 0; i < N; ++i)                                                                               \
      r[i] = a[i] OP s;                                                                                               \
    return r;                                                                                                         \
  }                                                                                                                   \
  template <typename T, IdComponent N>                                                                                \
  inline Vec<T, N> operator OP(T s, const Vec<T, N>& a)                                                               \
  {                                                                                                                   \
    Vec<T, N> r;                                                                                                      \
    for (IdComponent i = 0; i < N; ++i)                                                                               \
      r[i] = s OP a[i];                                                                                               \
    return r;                                                                                                         \
  }
ORACLE_VTKM_VEC_OP(+)
ORACLE_VTKM_VEC_OP(-)
ORACLE_VTKM_VEC_OP(*)
ORACLE_VTKM_VEC_OP(/)
#undef ORACLE_VTKM_VEC_OP

// Float32 vector with a Float64 scalar: computed in double, narrowed (VTK-m Types.h)
template <IdComponent N>
inline Vec<Float32, N> operator*(const Vec<Float32, N>& a, Float64 s)
{
  Vec<Float32, N> r;
  for (IdComponent i = 0; i < N; ++i)
    r[i] = static_cast<Float32>(static_cast<Float64>(a[i]) * s);
  return r;
}
template <IdComponent N>
inline Vec<Float32, N> operator*(Float64 s, const Vec<Float32, N>& a)
{
  Vec<Float32, N> r;
  for (IdComponent i = 0; i < N; ++i)
    r[i] = static_cast<Float32>(s * static_cast<Float64>(a[i]));
  return r;
}
template <IdComponent N>
inline Vec<Float32, N> operator/(const Vec<Float32, N>& a, Float64 s)
{
  Vec<Float32, N> r;
  for (IdComponent i = 0; i < N; ++i)
    r[i] = static_cast<Float32>(static_cast<Float64>(a[i]) / s);
  return r;
}

template <typename T>
inline T Dot(const Vec<T, 3>& a, const Vec<T, 3>& b)
{
  return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}
template <typename T>
inline T dot(const Vec<T, 3>& a, const Vec<T, 3>& b) // deprecated lower-case spelling (Surface.h:184)
{
  return Dot(a, b);
}
} // namespace vtkm
#endif
