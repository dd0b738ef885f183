// oracle/vtkm_min -- see vtkm/Types.h (TEST INFRASTRUCTURE, VTK-m stand-in)
#ifndef oracle_vtkm_min_VectorAnalysis_h
#define oracle_vtkm_min_VectorAnalysis_h
#include <vtkm/Math.h>
namespace vtkm
{
template <typename T>
inline T MagnitudeSquared(const Vec<T, 3>& v)
{
  return Dot(v, v);
}
template <typename T>
inline T Magnitude(const Vec<T, 3>& v)
{
  return Sqrt(MagnitudeSquared(v));
}
template <typename T>
inline T RMagnitude(const Vec<T, 3>& v)
{
  return RSqrt(MagnitudeSquared(v));
}
template <typename T>
inline void Normalize(Vec<T, 3>& v)
{
  v = v * RMagnitude(v);
}
template <typename T>
inline Vec<T, 3> Cross(const Vec<T, 3>& a, const Vec<T, 3>& b)
{
  return Vec<T, 3>(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
template <typename T>
inline Vec<T, 3> TriangleNormal(const Vec<T, 3>& a, const Vec<T, 3>& b, const Vec<T, 3>& c)
{
  return Cross(b - a, c - a);
}
} // namespace vtkm
#endif
