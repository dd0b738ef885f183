// oracle/vtkm_min -- see vtkm/Types.h (TEST INFRASTRUCTURE, VTK-m stand-in)
#ifndef oracle_vtkm_min_Math_h
#define oracle_vtkm_min_Math_h
#include <vtkm/Types.h>
namespace vtkm
{
inline Float64 Pi() { return 3.14159265358979323846264338327950288; }
inline Float32 Pi_180f() { return 0.01745329251994329547f; }
template <typename T>
inline T Epsilon();
template <>
inline Float32 Epsilon<Float32>() { return 1e-5f; }
template <>
inline Float64 Epsilon<Float64>() { return 1e-9; }
inline Float32 Min(Float32 a, Float32 b) { return std::fmin(a, b); }
inline Float32 Max(Float32 a, Float32 b) { return std::fmax(a, b); }
inline Float64 Min(Float64 a, Float64 b) { return std::fmin(a, b); }
inline Float64 Max(Float64 a, Float64 b) { return std::fmax(a, b); }
inline int Min(int a, int b) { return b < a ? b : a; }
inline int Max(int a, int b) { return a < b ? b : a; }
inline Id Min(Id a, Id b) { return b < a ? b : a; }
inline Id Max(Id a, Id b) { return a < b ? b : a; }
inline Float32 Abs(Float32 a) { return std::fabs(a); }
inline Float64 Abs(Float64 a) { return std::fabs(a); }
inline Float32 Sqrt(Float32 a) { return std::sqrt(a); }
inline Float64 Sqrt(Float64 a) { return std::sqrt(a); }
inline Float32 Pow(Float32 a, Float32 b) { return std::pow(a, b); }
inline Float64 Pow(Float64 a, Float64 b) { return std::pow(a, b); }
inline Float32 RSqrt(Float32 a) { return 1.0f / std::sqrt(a); }
inline Float64 RSqrt(Float64 a) { return 1.0 / std::sqrt(a); }
} // namespace vtkm
#endif
