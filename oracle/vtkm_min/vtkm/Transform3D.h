// oracle/vtkm_min -- see vtkm/Types.h (TEST INFRASTRUCTURE, VTK-m stand-in).  vtkm/Matrix.h + vtkm/Transform3D.h as
// VTK-m defines them: row-major Matrix, MatrixMultiply rows as left-to-right dot products (vtkm::Dot of Vec4 =
// ((a0b0 + a1b1) + a2b2) + a3b3), Transform3DRotate from the normalised axis and the angle in degrees, all in T.
#ifndef oracle_vtkm_min_Transform3D_h
#define oracle_vtkm_min_Transform3D_h
#include <vtkm/VectorAnalysis.h>
namespace vtkm
{
template <typename T, IdComponent R, IdComponent C>
struct Matrix
{
  T m[R][C];
  T& operator()(IdComponent r, IdComponent c) { return m[r][c]; }
  const T& operator()(IdComponent r, IdComponent c) const { return m[r][c]; }
};
template <typename T, IdComponent N>
inline Matrix<T, N, N> MatrixIdentity()
{
  Matrix<T, N, N> r;
  for (IdComponent i = 0; i < N; ++i)
    for (IdComponent j = 0; j < N; ++j)
      r(i, j) = (i == j) ? T(1) : T(0);
  return r;
}
template <typename T, IdComponent R, IdComponent C>
inline Matrix<T, C, R> MatrixTranspose(const Matrix<T, R, C>& a)
{
  Matrix<T, C, R> r;
  for (IdComponent i = 0; i < R; ++i)
    for (IdComponent j = 0; j < C; ++j)
      r(j, i) = a(i, j);
  return r;
}
template <typename T, IdComponent R, IdComponent K, IdComponent C>
inline Matrix<T, R, C> MatrixMultiply(const Matrix<T, R, K>& a, const Matrix<T, K, C>& b)
{
  Matrix<T, R, C> r;
  for (IdComponent i = 0; i < R; ++i)
    for (IdComponent j = 0; j < C; ++j)
    {
      T s = a(i, 0) * b(0, j);
      for (IdComponent k = 1; k < K; ++k)
        s = s + a(i, k) * b(k, j);
      r(i, j) = s;
    }
  return r;
}
template <typename T, IdComponent R, IdComponent C>
inline Vec<T, R> MatrixMultiply(const Matrix<T, R, C>& a, const Vec<T, C>& v)
{
  Vec<T, R> r;
  for (IdComponent i = 0; i < R; ++i)
  {
    T s = a(i, 0) * v[0];
    for (IdComponent k = 1; k < C; ++k)
      s = s + a(i, k) * v[k];
    r[i] = s;
  }
  return r;
}
template <typename T>
inline Matrix<T, 4, 4> Transform3DTranslate(const T& x, const T& y, const T& z)
{
  Matrix<T, 4, 4> r = MatrixIdentity<T, 4>();
  r(0, 3) = x;
  r(1, 3) = y;
  r(2, 3) = z;
  return r;
}
template <typename T>
inline Matrix<T, 4, 4> Transform3DRotate(T angleDegrees, T ax, T ay, T az)
{
  const T angleRadians = static_cast<T>(0.01745329251994329547) * angleDegrees; // Pi_180<T>()
  Vec<T, 3> n(ax, ay, az);
  Normalize(n);
  const T s = std::sin(angleRadians), c = std::cos(angleRadians);
  Matrix<T, 4, 4> r;
  r(0, 0) = n[0] * n[0] * (1 - c) + c;
  r(0, 1) = n[0] * n[1] * (1 - c) - n[2] * s;
  r(0, 2) = n[0] * n[2] * (1 - c) + n[1] * s;
  r(0, 3) = T(0);
  r(1, 0) = n[1] * n[0] * (1 - c) + n[2] * s;
  r(1, 1) = n[1] * n[1] * (1 - c) + c;
  r(1, 2) = n[1] * n[2] * (1 - c) - n[0] * s;
  r(1, 3) = T(0);
  r(2, 0) = n[2] * n[0] * (1 - c) - n[1] * s;
  r(2, 1) = n[2] * n[1] * (1 - c) + n[0] * s;
  r(2, 2) = n[2] * n[2] * (1 - c) + c;
  r(2, 3) = T(0);
  r(3, 0) = T(0);
  r(3, 1) = T(0);
  r(3, 2) = T(0);
  r(3, 3) = T(1);
  return r;
}
} // namespace vtkm
#endif
