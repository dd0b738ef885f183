/*
 * b2pt_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See b2pt_oracle.h.
 *
 * Plain C restatement of the reference hot path.  Build with
 *   gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -shared
 * (the reference's CMakeLists.txt:1-93 sets no -march/-mfma, i.e. an x86-64 baseline
 * build with no FMA contraction; -ffp-contract=off reproduces that on any host).
 *
 * PARITY STATUS: see b2pt_oracle.h -- pinned bit for bit to the reference's own worklets, Camera::RayGen and scene
 * builder through oracle/ref_harness.cxx + ref_scene.cxx; what stays "parity unpinned" is VTK-m's own math and
 * LinearBVH shape (VTK-m is absent, so the reference cannot be compiled as a whole here).
 * VTK-m math semantics assumed:
 *   Dot(a,b)      = (a0*b0 + a1*b1) + a2*b2
 *   Cross(a,b)    = (a1*b2 - a2*b1, a2*b0 - a0*b2, a0*b1 - a1*b0)   (plain, no FMA compensation)
 *   RSqrt(x)      = 1.0f / sqrtf(x)  (host form);  Normalize(v) = v * RSqrt(Dot(v,v))
 *   Magnitude(v)  = sqrtf(Dot(v,v)); RMagnitude = RSqrt(Dot(v,v)); Epsilon<Float32>() = 1e-5f
 *   Min/Max       = fminf/fmaxf;  Pi() is Float64.
 */
#include "b2pt_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.14159265358979323846
#define ORC_PI_180F 0.01745329251994329547f /* vtkm::Pi_180f() */
#define ORC_EPS 1e-5f                       /* vtkm::Epsilon<Float32>() */
#define ORC_GOLDEN 0x9E3779B9u

typedef struct
{
  float x, y, z;
} v3;

static inline v3 V(float x, float y, float z)
{
  v3 r = { x, y, z };
  return r;
}
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline float vdot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline v3 vcross(v3 a, v3 b)
{
  return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float rsqrt_host(float x) { return 1.0f / sqrtf(x); }
static inline float vrmag(v3 a) { return rsqrt_host(vdot(a, a)); }
static inline float vmag(v3 a) { return sqrtf(vdot(a, a)); }
static inline v3 vnormalize(v3 a) { return vscale(a, vrmag(a)); } /* vtkm::Normalize */
static inline v3 unit_vector(v3 a) { return vscale(a, vrmag(a)); } /* vec3.h:38-42 */
static inline v3 ld3(const float* p) { return V(p[0], p[1], p[2]); }
static inline void st3(float* p, v3 a)
{
  p[0] = a.x;
  p[1] = a.y;
  p[2] = a.z;
}
static inline v3 de_nan3(v3 c) /* PdfWorklet.h:39-45 */
{
  if (!(c.x == c.x))
    c.x = 0;
  if (!(c.y == c.y))
    c.y = 0;
  if (!(c.z == c.z))
    c.z = 0;
  return c;
}

/* ------------------------------------------------------------------------------------------ RNG */
/* wangXor.h:30-38 */
uint32_t orc_wang32(uint32_t* seed)
{
  uint32_t s = *seed;
  s = (s ^ 61u) ^ (s >> 16);
  s *= 9u;
  s = s ^ (s >> 4);
  s *= 0x27d4eb2du;
  s = s ^ (s >> 15);
  *seed = s;
  return s;
}
/* wangXor.h:55-59 */
float orc_randf(uint32_t* seed)
{
  uint32_t t = orc_wang32(seed);
  return (float)t / 4294967295.f;
}
/* MapperPathTracer.cxx:60-75 (the CopyIf predicate; seeds end up = counting index) */
uint32_t orc_wang_init(uint32_t x)
{
  uint32_t idx = x;
  uint32_t val = orc_wang32(&idx);
  orc_wang32(&val);
  orc_wang32(&val);
  orc_wang32(&val);
  return val;
}

/* -------------------------------------------------------------------------------------- scene */
/* vtkm::Transform3DRotate(angle, axis) for float, then transposed and pre-multiplied by a translation,
 * CornellBox.cpp:10-35.  Matrix-vector product rows use the 4-term left-to-right Dot. */
static void cornell_invert(v3* p4)
{
  const float angleDeg = -15.f;
  const float ang = ORC_PI_180F * angleDeg;
  const float ax = 0.f, ay = 1.f, az = 0.f; /* already unit */
  const float s = sinf(ang), c = cosf(ang);
  float R[4][4];
  memset(R, 0, sizeof(R));
  R[0][0] = ax * ax * (1 - c) + c;
  R[0][1] = ax * ay * (1 - c) - az * s;
  R[0][2] = ax * az * (1 - c) + ay * s;
  R[1][0] = ay * ax * (1 - c) + az * s;
  R[1][1] = ay * ay * (1 - c) + c;
  R[1][2] = ay * az * (1 - c) - ax * s;
  R[2][0] = az * ax * (1 - c) - ay * s;
  R[2][1] = az * ay * (1 - c) + ax * s;
  R[2][2] = az * az * (1 - c) + c;
  R[3][3] = 1.f;
  float Rt[4][4], T[4][4], M[4][4];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      Rt[i][j] = R[j][i];
  memset(T, 0, sizeof(T));
  T[0][0] = T[1][1] = T[2][2] = T[3][3] = 1.f;
  T[0][3] = 265.f;
  T[1][3] = 0.f;
  T[2][3] = 295.f;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
    {
      float sum = 0.f;
      for (int k = 0; k < 4; k++)
        sum += T[i][k] * Rt[k][j];
      M[i][j] = sum;
    }
  for (int k = 0; k < 4; k++)
  {
    float in[4] = { p4[k].x, p4[k].y, p4[k].z, 1.f };
    float out[3];
    for (int r = 0; r < 3; r++)
      out[r] = ((M[r][0] * in[0] + M[r][1] * in[1]) + M[r][2] * in[2]) + M[r][3] * in[3];
    p4[k] = V(out[0], out[1], out[2]);
  }
}

typedef struct
{
  float* pts;
  int64_t nPts;
  int64_t* quadIds;
  int64_t nQuads;
  int64_t* matQ;
  int64_t* texQ;
} scene_builder;

/* CornellBox.cpp:37-61: pts[k]/555.0 is a Float64 division narrowed to Float32 */
static void sb_quad(scene_builder* sb, const v3* p4, int divide, int m, int t)
{
  int64_t q = sb->nQuads++;
  sb->quadIds[5 * q] = q < 12 ? q : q + 1; /* cell id: the sphere VERTEX cell sits at index 12 */
  for (int k = 0; k < 4; k++)
  {
    int64_t pi = sb->nPts++;
    float x = p4[k].x, y = p4[k].y, z = p4[k].z;
    if (divide)
    {
      x = (float)((double)x / 555.0);
      y = (float)((double)y / 555.0);
      z = (float)((double)z / 555.0);
    }
    sb->pts[3 * pi] = x;
    sb->pts[3 * pi + 1] = y;
    sb->pts[3 * pi + 2] = z;
    sb->quadIds[5 * q + 1 + k] = pi;
  }
  sb->matQ[q] = m;
  sb->texQ[q] = t;
}
/* CornellBox.cpp:63-139: five faces; faces 4,5 repeat faces 1,2 */
static void sb_box(scene_builder* sb, v3 n, v3 f, int m, int t)
{
  v3 p[4];
  p[0] = V(n.x, n.y, n.z), p[1] = V(f.x, n.y, n.z), p[2] = V(f.x, f.y, n.z), p[3] = V(n.x, f.y, n.z);
  sb_quad(sb, p, 1, m, t);
  p[0] = V(n.x, n.y, f.z), p[1] = V(f.x, n.y, f.z), p[2] = V(f.x, f.y, f.z), p[3] = V(n.x, f.y, f.z);
  sb_quad(sb, p, 1, m, t);
  p[0] = V(n.x, f.y, n.z), p[1] = V(f.x, f.y, n.z), p[2] = V(f.x, f.y, f.z), p[3] = V(n.x, f.y, f.z);
  sb_quad(sb, p, 1, m, t);
  p[0] = V(n.x, n.y, n.z), p[1] = V(f.x, n.y, n.z), p[2] = V(f.x, f.y, n.z), p[3] = V(n.x, f.y, n.z);
  sb_quad(sb, p, 1, m, t);
  p[0] = V(n.x, n.y, f.z), p[1] = V(f.x, n.y, f.z), p[2] = V(f.x, f.y, f.z), p[3] = V(n.x, f.y, f.z);
  sb_quad(sb, p, 1, m, t);
}

/* CornellBox.cpp:141-418 */
int orc_cornell_scene(float* pts, int64_t* quadIds, int64_t* sphPt, float* sphR, int64_t* matIdxQ, int64_t* texIdxQ,
                      int64_t* matIdxS, int64_t* texIdxS, int* matType, int* texType, float* tex)
{
  static const float texv[12] = { 0.65f, 0.05f, 0.05f, 0.73f, 0.73f, 0.73f, 0.12f, 0.45f, 0.15f, 15.f, 15.f, 15.f };
  static const int mt[5] = { 0, 0, 0, 1, 2 };
  static const int tt[5] = { 0, 1, 2, 3, 0 };
  memcpy(tex, texv, sizeof(texv));
  memcpy(matType, mt, sizeof(mt));
  memcpy(texType, tt, sizeof(tt));
  scene_builder sb = { pts, 0, quadIds, 0, matIdxQ, texIdxQ };
  v3 p[4];
  /* :175-186 green wall x=555 */
  p[0] = V(555, 0, 0), p[1] = V(555, 555, 0), p[2] = V(555, 555, 555), p[3] = V(555, 0, 555);
  sb_quad(&sb, p, 1, 2, 2);
  /* :189-200 red wall x=0 (first vertex is vec3(0,0,0) undivided -- identical value) */
  p[0] = V(0, 0, 0), p[1] = V(0, 555, 0), p[2] = V(0, 555, 555), p[3] = V(0, 0, 555);
  sb_quad(&sb, p, 1, 0, 0);
  /* :204-215 light */
  p[0] = V(213, 554, 227), p[1] = V(343, 554, 227), p[2] = V(343, 554, 332), p[3] = V(213, 554, 332);
  sb_quad(&sb, p, 1, 3, 3);
  /* :218-229 ceiling */
  p[0] = V(0, 555, 0), p[1] = V(555, 555, 0), p[2] = V(555, 555, 555), p[3] = V(0, 555, 555);
  sb_quad(&sb, p, 1, 1, 1);
  /* :232-243 floor */
  p[0] = V(0, 0, 0), p[1] = V(555, 0, 0), p[2] = V(555, 0, 555), p[3] = V(0, 0, 555);
  sb_quad(&sb, p, 1, 1, 1);
  /* :247-258 back wall */
  p[0] = V(0, 0, 555), p[1] = V(555, 0, 555), p[2] = V(555, 555, 555), p[3] = V(0, 555, 555);
  sb_quad(&sb, p, 1, 1, 1);
  /* :264-353 tall box, six faces through invert() */
  p[0] = V(0, 0, 165), p[1] = V(165, 0, 165), p[2] = V(165, 330, 165), p[3] = V(0, 330, 165);
  cornell_invert(p);
  sb_quad(&sb, p, 1, 1, 1);
  p[0] = V(0, 0, 0), p[1] = V(165, 0, 0), p[2] = V(165, 330, 0), p[3] = V(0, 330, 0);
  cornell_invert(p);
  sb_quad(&sb, p, 1, 1, 1);
  p[0] = V(165, 0, 0), p[1] = V(165, 330, 0), p[2] = V(165, 330, 165), p[3] = V(165, 0, 165);
  cornell_invert(p);
  sb_quad(&sb, p, 1, 1, 1);
  p[0] = V(0, 0, 0), p[1] = V(0, 330, 0), p[2] = V(0, 330, 165), p[3] = V(0, 0, 165);
  cornell_invert(p);
  sb_quad(&sb, p, 1, 1, 1);
  p[0] = V(0, 333, 0), p[1] = V(165, 330, 0), p[2] = V(165, 330, 165), p[3] = V(0, 330, 165); /* :327 typo kept */
  cornell_invert(p);
  sb_quad(&sb, p, 1, 1, 1);
  p[0] = V(0, 0, 0), p[1] = V(165, 0, 0), p[2] = V(165, 0, 165), p[3] = V(0, 0, 165);
  cornell_invert(p);
  sb_quad(&sb, p, 1, 1, 1);
  /* :357-365 sphere centre, point 48 */
  {
    int64_t pi = sb.nPts++;
    pts[3 * pi] = (float)(-335.0 / 555.0);
    pts[3 * pi + 1] = (float)(90.0 / 555.0);
    pts[3 * pi + 2] = (float)(290.0 / 555.0);
    sphPt[0] = pi;
    sphR[0] = (float)(90 / 555.0); /* MapperPathTracer.cxx:182 */
    matIdxS[0] = 4;
    texIdxS[0] = 0;
  }
  /* :368-375 */
  sb_box(&sb, V(135.f - 90.f, 0, 290.f - 90.f), V(135.f + 90.f, 180, 290.f + 90.f), 1, 1);
  /* :379-386 */
  sb_box(&sb, V(50, 0, 50), V(450, 100, 100), 1, 1);
  return (sb.nPts == 89 && sb.nQuads == 22) ? 0 : -1;
}

/* ------------------------------------------------------------------------------------- camera */
/* Camera.cxx:438-476 (RayGen ctor), :913-914 (Look), :803-811 (SetUp normalises) */
void orc_camera_basis(const orc_camera* cam, float* nlook3, float* dx3, float* dy3)
{
  v3 look = vnormalize(vsub(ld3(cam->lookAt), ld3(cam->pos)));
  v3 up = ld3(cam->up);
  if (!(up.x == 0.f && up.y == 1.f && up.z == 0.f))
    up = vnormalize(up);
  float thx = tanf((cam->fovDeg * ORC_PI_180F) * .5f);
  float thy = tanf((cam->fovDeg * ORC_PI_180F) * .5f); /* fovX = fovY, Camera.cxx:936-938 */
  v3 u = vnormalize(vcross(look, up));
  v3 v = vnormalize(vcross(u, look));
  st3(dx3, vscale(u, 2 * thx / (float)cam->W));
  st3(dy3, vscale(v, 2 * thy / (float)cam->H));
  st3(nlook3, vnormalize(look));
}

typedef struct
{
  v3 nlook, dx, dy, pos;
  int W, H;
} cam_basis;

static cam_basis make_basis(const orc_camera* cam)
{
  cam_basis b;
  float a[3], c[3], d[3];
  orc_camera_basis(cam, a, c, d);
  b.nlook = ld3(a);
  b.dx = ld3(c);
  b.dy = ld3(d);
  b.pos = ld3(cam->pos);
  b.W = cam->W;
  b.H = cam->H;
  return b;
}

/* Camera.cxx:483-524 */
static inline v3 raygen(const cam_basis* b, int64_t idx, uint32_t* seed)
{
  int i = (int)((int32_t)idx % b->W);
  int j = (int)((int32_t)idx / b->W);
  float ru = orc_randf(seed);
  float rv = orc_randf(seed);
  float sx = (2.f * ((float)i + (1.f - ru)) - (float)b->W) / 2.0f;
  float sy = (2.f * ((float)j + (rv)) - (float)b->H) / 2.0f;
  v3 d = vadd(vadd(b->nlook, vscale(b->dx, sx)), vscale(b->dy, sy));
  if (d.x == 0.f)
    d.x += 0.0000001f;
  if (d.y == 0.f)
    d.y += 0.0000001f;
  if (d.z == 0.f)
    d.z += 0.0000001f;
  float dot = vdot(d, d);
  float m = sqrtf(dot);
  return V(d.x / m, d.y / m, d.z / m);
}
void orc_raygen(const orc_camera* cam, int64_t idx, uint32_t* seed, float* dir3)
{
  cam_basis b = make_basis(cam);
  st3(dir3, raygen(&b, idx, seed));
}

/* --------------------------------------------------------------------------------- primitives */
/* Surface.h:30-161 (Lagae-Dutre) */
static int quad_hit(v3 ro, v3 rd, v3 v00, v3 v10, v3 v11, v3 v01, float* u, float* v, float* t)
{
  v3 E03 = vsub(v01, v00);
  v3 P = vcross(rd, E03);
  v3 E01 = vsub(v10, v00);
  float det = vdot(E01, P);
  if (fabsf(det) < ORC_EPS)
    return 0;
  float inv_det = 1.0f / det;
  v3 T = vsub(ro, v00);
  float alpha = vdot(T, P) * inv_det;
  if (alpha < 0.0f)
    return 0;
  v3 Q = vcross(T, E01);
  float beta = vdot(rd, Q) * inv_det;
  if (beta < 0.0f)
    return 0;
  if ((alpha + beta) > 1.0f)
  {
    v3 E23 = vsub(v01, v11);
    v3 E21 = vsub(v10, v11);
    v3 Pp = vcross(rd, E21);
    float detp = vdot(E23, Pp);
    if (fabsf(detp) < ORC_EPS)
      return 0;
    float inv_detp = 1.0f / detp;
    v3 Tp = vsub(ro, v11);
    float alphap = vdot(Tp, Pp) * inv_detp;
    if (alphap < 0.0f)
      return 0;
    v3 Qp = vcross(Tp, E23);
    float betap = vdot(rd, Qp) * inv_detp;
    if (betap < 0.0f)
      return 0;
  }
  *t = vdot(E03, Q) * inv_det;
  if (*t < 0.0f)
    return 0;
  /* bilinear u,v (Surface.h:106-158): computed by the reference, never consumed downstream */
  float alpha_11, beta_11;
  v3 E02 = vsub(v11, v00);
  v3 n = vcross(E01, E02);
  float anx = fabsf(n.x), any = fabsf(n.y), anz = fabsf(n.z);
  if ((anx >= any) && (anx >= anz))
  {
    alpha_11 = ((E02.y * E03.z) - (E02.z * E03.y)) / n.x;
    beta_11 = ((E01.y * E02.z) - (E01.z * E02.y)) / n.x;
  }
  else if ((any >= anx) && (any >= anz))
  {
    alpha_11 = ((E02.z * E03.x) - (E02.x * E03.z)) / n.y;
    beta_11 = ((E01.z * E02.x) - (E01.x * E02.z)) / n.y;
  }
  else
  {
    alpha_11 = ((E02.x * E03.y) - (E02.y * E03.x)) / n.z;
    beta_11 = ((E01.x * E02.y) - (E01.y * E02.x)) / n.z;
  }
  if (fabsf(alpha_11 - 1.0f) < ORC_EPS)
  {
    *u = alpha;
    if (fabsf(beta_11 - 1.0f) < ORC_EPS)
      *v = beta;
    else
      *v = beta / ((*u * (beta_11 - 1.0f)) + 1.0f);
  }
  else if (fabs((double)beta_11 - 1.0) < (double)ORC_EPS)
  {
    *v = beta;
    *u = alpha / ((*v * (alpha_11 - 1.0f)) + 1.0f);
  }
  else
  {
    float A = 1.0f - beta_11;
    float B = (alpha * (beta_11 - 1.0f)) - (beta * (alpha_11 - 1.0f)) - 1.0f;
    float C = alpha;
    float D = (B * B) - (4.0f * A * C);
    float QQ = -0.5f * (B + ((B < 0.0f ? -1.0f : 1.0f) * sqrtf(D)));
    *u = QQ / A;
    if ((*u < 0.0f) || (*u > 1.0f))
      *u = C / QQ;
    *v = beta / ((*u * (beta_11 - 1.0f)) + 1.0f);
  }
  return 1;
}
int orc_quad_hit(const float* o3, const float* d3, const float* v00, const float* v10, const float* v11,
                 const float* v01, float* u, float* v, float* t)
{
  return quad_hit(ld3(o3), ld3(d3), ld3(v00), ld3(v10), ld3(v11), ld3(v01), u, v, t);
}

/* hrec layout Record.h:4: U,V,T,Nx,Ny,Nz,Px,Py,Pz */
enum
{
  HR_U,
  HR_V,
  HR_T,
  HR_NX,
  HR_NY,
  HR_NZ,
  HR_PX,
  HR_PY,
  HR_PZ
};

/* Surface.h:163-199 */
static int quad_intersect(v3 ro, v3 rd, float* rec9, float tmin, float tmax, v3 q, v3 r, v3 s, v3 t)
{
  float u = 0, v = 0, tt = 0;
  int h = quad_hit(ro, rd, q, r, s, t, &u, &v, &tt);
  rec9[HR_U] = u;
  rec9[HR_V] = v;
  rec9[HR_T] = tt;
  h = h && (tt < tmax) && (tt > tmin);
  if (h)
  {
    v3 normal = vcross(vsub(r, q), vsub(s, q)); /* vtkm::TriangleNormal */
    normal = vnormalize(normal);
    if (vdot(normal, rd) > 0.f)
      normal = vneg(normal);
    v3 p = vadd(ro, vscale(rd, tt));
    rec9[HR_PX] = p.x;
    rec9[HR_PY] = p.y;
    rec9[HR_PZ] = p.z;
    rec9[HR_NX] = normal.x;
    rec9[HR_NY] = normal.y;
    rec9[HR_NZ] = normal.z;
  }
  return h;
}

/* Surface.h:319-367 */
static int sphere_hit(v3 ro, v3 rd, float* rec9, float tmin, float tmax, v3 center, float radius)
{
  v3 oc = vsub(ro, center);
  float a = vdot(rd, rd);
  float b = vdot(oc, rd);
  float c = vdot(oc, oc) - radius * radius;
  float discriminant = b * b - a * c;
  if (discriminant > 0)
  {
    for (int root = 0; root < 2; root++)
    {
      float temp = root == 0 ? (-b - sqrtf(b * b - a * c)) / a : (-b + sqrtf(b * b - a * c)) / a;
      if (temp < tmax && temp > tmin)
      {
        rec9[HR_T] = temp;
        v3 p = vadd(ro, vscale(rd, temp));
        rec9[HR_PX] = p.x;
        rec9[HR_PY] = p.y;
        rec9[HR_PZ] = p.z;
        v3 pn = V((p.x - center.x) / radius, (p.y - center.y) / radius, (p.z - center.z) / radius);
        /* get_sphere_uv, Surface.h:310-315 (unused downstream) */
        float phi = atan2f(pn.z, pn.x);
        float theta = asinf(pn.y);
        rec9[HR_U] = (float)(1 - ((double)phi + ORC_PI) / (2 * ORC_PI));
        rec9[HR_V] = (float)(((double)theta + ORC_PI / 2) / ORC_PI);
        rec9[HR_NX] = pn.x;
        rec9[HR_NY] = pn.y;
        rec9[HR_NZ] = pn.z;
        return 1;
      }
    }
  }
  return 0;
}

/* BVHTraverser.h:35-79 applied to a single box; the leaf test of the reference's BVH (one primitive per leaf) */
static inline float rcp_safe(float f) { return 1.0f / ((fabsf(f) < 1e-8f) ? 1e-8f : f); }
static int aabb_gate(const float* bb6, v3 ro, v3 rd, float tmin, float closest)
{
  v3 inv = V(rcp_safe(rd.x), rcp_safe(rd.y), rcp_safe(rd.z));
  v3 od = V(ro.x * inv.x, ro.y * inv.y, ro.z * inv.z);
  float xmin0 = bb6[0] * inv.x - od.x, ymin0 = bb6[1] * inv.y - od.y, zmin0 = bb6[2] * inv.z - od.z;
  float xmax0 = bb6[3] * inv.x - od.x, ymax0 = bb6[4] * inv.y - od.y, zmax0 = bb6[5] * inv.z - od.z;
  float min0 = fmaxf(fmaxf(fmaxf(fminf(ymin0, ymax0), fminf(xmin0, xmax0)), fminf(zmin0, zmax0)), tmin);
  float max0 = fminf(fminf(fminf(fmaxf(ymin0, ymax0), fmaxf(xmin0, xmax0)), fmaxf(zmin0, zmax0)), closest);
  return max0 >= min0;
}
/* AABBSurface.h:24-78 */
static void quad_aabb(v3 q, v3 r, v3 s, v3 t, float* bb6)
{
  float xmin = fminf(fminf(fminf(q.x, r.x), s.x), t.x), xmax = fmaxf(fmaxf(fmaxf(q.x, r.x), s.x), t.x);
  float ymin = fminf(fminf(fminf(q.y, r.y), s.y), t.y), ymax = fmaxf(fmaxf(fmaxf(q.y, r.y), s.y), t.y);
  float zmin = fminf(fminf(fminf(q.z, r.z), s.z), t.z), zmax = fmaxf(fmaxf(fmaxf(q.z, r.z), s.z), t.z);
  float xe = fmaxf(1e-6f, 1.0e-4f * (xmax - xmin));
  float ye = fmaxf(1e-6f, 1.0e-4f * (ymax - ymin));
  float ze = fmaxf(1e-6f, 1.0e-4f * (zmax - zmin));
  bb6[0] = xmin - xe, bb6[1] = ymin - ye, bb6[2] = zmin - ze;
  bb6[3] = xmax + xe, bb6[4] = ymax + ye, bb6[5] = zmax + ze;
}
/* AABBSurface.h:98-172 (no padding) */
static void sphere_aabb(v3 c, float r, float* bb6)
{
  bb6[0] = fminf(c.x + r, c.x - r), bb6[3] = fmaxf(c.x + r, c.x - r);
  bb6[1] = fminf(c.y + r, c.y - r), bb6[4] = fmaxf(c.y + r, c.y - r);
  bb6[2] = fminf(c.z + r, c.z - r), bb6[5] = fmaxf(c.z + r, c.z - r);
  /* the reference also mins/maxes the un-offset coordinates of the other axes: same result */
}

/* planar within 1e-5 of the largest edge: both out-of-triangle vertices' distances to the (q,r,s) plane */
static int quad_is_planar(v3 q, v3 r, v3 s, v3 t)
{
  v3 n = vnormalize(vcross(vsub(r, q), vsub(s, q)));
  v3 e1 = vsub(r, q), e2 = vsub(s, q), e3 = vsub(t, q);
  float scale = fmaxf(fmaxf(vmag(e1), vmag(e2)), vmag(e3));
  float dev = fmaxf(fabsf(vdot(n, e2)), fabsf(vdot(n, e3)));
  return dev <= 1e-5f * scale;
}

static inline void quad_pts(const orc_scene* sc, int64_t q, v3* a, v3* b, v3* c, v3* d)
{
  const int64_t* id = sc->quadIds + 5 * q;
  *a = ld3(sc->pts + 3 * id[1]);
  *b = ld3(sc->pts + 3 * id[2]);
  *c = ld3(sc->pts + 3 * id[3]);
  *d = ld3(sc->pts + 3 * id[4]);
}

/* Closest hit: quads (QuadIntersector.cxx:59-71 -> Surface.h:201-254) then spheres continuing from the
 * quads' tmax (SphereIntersector.cxx:88-100 -> Surface.h:369-409).  Primitives are visited in index order.
 * The reference reaches a primitive only through BVHTraverser (one primitive per LinearBVH leaf), i.e. only
 * if the ray passes the slab test of that primitive's own AABB (ancestor boxes are supersets and pass
 * whenever the leaf box does).  That gate is part of the acceptance rule: the Lagae-Dutre test returns
 * spurious far hits on the NON-PLANAR tall-box top face (CornellBox.cpp:327 typo vertex), which the leaf
 * box culls.  For planar quads the gate never rejects an accepted hit.  Tree topology otherwise only
 * changes the order of exact-t ties (strict t<tmax keeps the first tested primitive). */
static int64_t closest_hit(const orc_scene* sc, v3 ro, v3 rd, float tmin, float tmax, int flags, float* hrec9,
                           int* hid2)
{
  int64_t prim = -1;
  float closest = tmax;
  float tmp[9];
  /* Two passes over the quads: planar quads first, then non-planar ones.  In the reference's near-child-first
   * traversal a leaf is visited only if its box entry does not lie beyond the closest hit found so far; for
   * legitimate (on-surface) hits that never matters, but the spurious off-surface hits of a non-planar quad
   * are accepted only if nothing closer than the box entry was found among the other primitives -- which
   * this order reproduces independently of tree topology (DESIGN.md "leaf-box gate"). */
  for (int pass = 0; pass < 2; pass++)
    for (int64_t q = 0; q < sc->nQuads; q++)
    {
      v3 a, b, c, d;
      quad_pts(sc, q, &a, &b, &c, &d);
      if (!(flags & ORC_FLAG_NO_AABB_GATE))
      {
        if (quad_is_planar(a, b, c, d) != (pass == 0))
          continue;
        float bb[6];
        quad_aabb(a, b, c, d, bb);
        if (!aabb_gate(bb, ro, rd, tmin, closest))
          continue;
      }
      else if (pass == 1)
        continue;
      if (quad_intersect(ro, rd, tmp, tmin, closest, a, b, c, d))
      {
        memcpy(hrec9, tmp, sizeof(tmp));
        closest = tmp[HR_T];
        hid2[0] = (int)sc->matIdxQ[q];
        hid2[1] = (int)sc->texIdxQ[q];
        prim = q;
      }
    }
  for (int64_t s = 0; s < sc->nSph; s++)
  {
    v3 c = ld3(sc->pts + 3 * sc->sphPt[s]);
    float r = sc->sphR[s];
    if (!(flags & ORC_FLAG_NO_AABB_GATE))
    {
      /* The sphere's leaf box is tested against tmax, not against the running closest distance: for small
       * spheres far from the ray origin the analytic test below cancels badly and reports "hits" slightly outside
       * the surface, whose t may lie before the box entry; with the running distance the winner among two such
       * near-coincident spheres would depend on the order of the primitives (for the reference: on VTK-m's tree
       * shape).  Against tmax the result is order independent.  The reference itself cannot hold more than one
       * sphere (SphereExtractor.cxx:108-111); for its single Cornell sphere both forms agree. */
      float bb[6];
      sphere_aabb(c, r, bb);
      if (!aabb_gate(bb, ro, rd, tmin, tmax))
        continue;
    }
    if (sphere_hit(ro, rd, tmp, tmin, closest, c, r))
    {
      memcpy(hrec9, tmp, sizeof(tmp));
      closest = tmp[HR_T];
      hid2[0] = (int)sc->matIdxS[s];
      hid2[1] = (int)sc->texIdxS[s];
      prim = sc->nQuads + s;
    }
  }
  return prim;
}
int64_t orc_closest_hit(const orc_scene* sc, const float* o3, const float* d3, float tmin, float tmax, int flags,
                        float* hrec9, int* hid2)
{
  return closest_hit(sc, ld3(o3), ld3(d3), tmin, tmax, flags, hrec9, hid2);
}

/* --------------------------------------------------------------------------- sampling and pdfs */
typedef struct
{
  v3 u, v, w;
} onb_t;
/* onb.h:34-45 */
static inline onb_t onb_from_w(v3 n)
{
  onb_t o;
  o.w = unit_vector(n);
  v3 a;
  if ((double)fabsf(o.w.x) > 0.9)
    a = V(0, 1, 0);
  else
    a = V(1, 0, 0);
  o.v = unit_vector(vcross(o.w, a));
  o.u = vcross(o.w, o.v);
  return o;
}
/* onb.h:30-31 */
static inline v3 onb_local(const onb_t* o, v3 a)
{
  return vadd(vadd(vscale(o->u, a.x), vscale(o->v, a.y)), vscale(o->w, a.z));
}
/* PdfWorklet.h:47-53 (note the 2*sqrt(r2): kept) */
static inline v3 random_cosine_direction(float r1, float r2)
{
  float z = sqrtf(1 - r2);
  float phi = (float)(2 * ORC_PI * (double)r1);
  float x = cosf(phi) * 2 * sqrtf(r2);
  float y = sinf(phi) * 2 * sqrtf(r2);
  return V(x, y, z);
}
/* PdfWorklet.h:157-165 */
static inline v3 random_to_sphere(float radius, float distance_squared, float r1, float r2)
{
  float z = 1 + r2 * (sqrtf(1 - radius * radius / distance_squared) - 1);
  float phi = (float)(2 * ORC_PI * (double)r1);
  float x = cosf(phi) * sqrtf(1 - z * z);
  float y = sinf(phi) * sqrtf(1 - z * z);
  return V(x, y, z);
}

/* PdfWorklet.h:230-248 */
static float quad_pdf_value(v3 o, v3 v, v3 q, v3 r, v3 s, v3 t)
{
  float rec[9];
  if (quad_intersect(o, v, rec, 0.001f, FLT_MAX, q, r, s, t))
  {
    float qr = vmag(vsub(r, q));
    float qt = vmag(vsub(t, q));
    float area = qr * qt;
    float rect = rec[HR_T];
    float distance_squared = rect * rect * vdot(v, v);
    v3 n = V(rec[HR_NX], rec[HR_NY], rec[HR_NZ]);
    float cosine = fabsf(vdot(v, n) * vrmag(v));
    return distance_squared / (cosine * area);
  }
  return 0;
}
float orc_quad_pdf_value(const float* o3, const float* v3_, const float* q, const float* r, const float* s,
                         const float* t)
{
  return quad_pdf_value(ld3(o3), ld3(v3_), ld3(q), ld3(r), ld3(s), ld3(t));
}
/* PdfWorklet.h:333-347 */
static float sphere_pdf_value(v3 o, v3 v, v3 center, float radius)
{
  float rec[9];
  if (sphere_hit(o, v, rec, 0.001f, FLT_MAX, center, radius))
  {
    v3 co = vsub(center, o);
    float cos_theta_max = sqrtf(1 - radius * radius / vdot(co, co));
    float solid_angle = (float)(2 * ORC_PI * (double)(1 - cos_theta_max));
    return 1 / solid_angle;
  }
  return 0;
}
float orc_sphere_pdf_value(const float* o3, const float* v3_, const float* c3, float radius)
{
  return sphere_pdf_value(ld3(o3), ld3(v3_), ld3(c3), radius);
}

/* EmitWorklet.h:152-226 (DielectricWorklet::schlick/refract/reflect/scatter).  Returns next direction. */
static v3 dielectric_scatter(v3 direction, v3 n, float ref_idx, float rnd)
{
  float dn = vdot(direction, n);
  v3 reflected = vsub(direction, vscale(n, 2 * dn)); /* v - 2*dot(v,n)*n */
  v3 outward_normal;
  float ni_over_nt, cosine, reflect_prob;
  if (dn > 0)
  {
    outward_normal = vneg(n);
    ni_over_nt = ref_idx;
    cosine = ref_idx * dn * vrmag(direction);
  }
  else
  {
    outward_normal = n;
    ni_over_nt = (float)(1.0 / (double)ref_idx);
    cosine = -dn * vrmag(direction);
  }
  v3 refracted = V(0, 0, 0);
  int did_refract = 0;
  {
    v3 uv = unit_vector(direction);
    float dt = vdot(uv, outward_normal);
    float discriminant = (float)(1.0 - (double)(ni_over_nt * ni_over_nt * (1 - dt * dt)));
    if (discriminant > 0)
    {
      refracted = vsub(vscale(vsub(uv, vscale(outward_normal, dt)), ni_over_nt),
                       vscale(outward_normal, sqrtf(discriminant)));
      did_refract = 1;
    }
  }
  if (did_refract)
  {
    float r0 = (1 - ref_idx) / (1 + ref_idx);
    r0 = r0 * r0;
    reflect_prob = (float)((double)r0 + (double)(1 - r0) * pow((double)(1 - cosine), 5.0));
  }
  else
    reflect_prob = 1.0f;
  return (rnd < reflect_prob) ? reflected : refracted;
}

/* ScatterWorklet.h:20-28 (cosine_pdf::value) and :52-58 (scattering_pdf), mixture :95-110 */
static inline void lambert_attenuation(v3 n, v3 g, float sum_value, v3 albedo, v3* atten)
{
  onb_t uvw = onb_from_w(n);
  float cosine = vdot(unit_vector(g), uvw.w);
  float value = (cosine > 0) ? (float)((double)cosine / ORC_PI) : 0.f;
  double pdf_val = 0.5 * (double)sum_value + 0.5 * (double)value;
  float c2 = vdot(n, unit_vector(g));
  float spdf = (c2 < 0) ? 0.f : (float)((double)c2 / ORC_PI);
  double sctr = (double)spdf / pdf_val;
  atten->x = (float)((double)albedo.x * sctr);
  atten->y = (float)((double)albedo.y * sctr);
  atten->z = (float)((double)albedo.z * sctr);
}

/* ------------------------------------------------------------------------ faithful stage form */
/* Status bits, Ray.h:32-39 / SURVEY A.2 */
#define BIT_FIN 1
#define BIT_HIT 2
#define BIT_SCT 3
#define BIT_SPEC 4

typedef struct
{
  int64_t n; /* slots */
  int depthCount;
  float *ox, *oy, *oz, *dx, *dy, *dz;
  float* hrec; /* 9*n, field-major [f*n+i] */
  int* hid;    /* 2*n */
  float* srec; /* 9*n: Ox,Oy,Oz,Dx,Dy,Dz,Ax,Ay,Az */
  unsigned char* status;
  int* which;
  float* gen;    /* 3*n */
  float* sum;    /* n */
  float* tmin;   /* n */
  float* atten;  /* 3*depthCount*n : [c][d*n+i] */
  float* emit;   /* same */
  float* sumtot; /* 3*n */
} State;

static int state_alloc(State* S, int64_t n, int depthCount)
{
  memset(S, 0, sizeof(*S));
  S->n = n;
  S->depthCount = depthCount;
#define AL(field, type, count)                                                                                        \
  S->field = (type*)calloc((size_t)(count), sizeof(type));                                                            \
  if (!S->field)                                                                                                       \
    return -1;
  AL(ox, float, n) AL(oy, float, n) AL(oz, float, n) AL(dx, float, n) AL(dy, float, n) AL(dz, float, n);
  AL(hrec, float, 9 * n) AL(hid, int, 2 * n) AL(srec, float, 9 * n) AL(status, unsigned char, n);
  AL(which, int, n) AL(gen, float, 3 * n) AL(sum, float, n) AL(tmin, float, n);
  AL(atten, float, 3 * (int64_t)depthCount * n) AL(emit, float, 3 * (int64_t)depthCount * n);
  AL(sumtot, float, 3 * n);
#undef AL
  return 0;
}
static void state_free(State* S)
{
  free(S->ox), free(S->oy), free(S->oz), free(S->dx), free(S->dy), free(S->dz);
  free(S->hrec), free(S->hid), free(S->srec), free(S->status), free(S->which), free(S->gen);
  free(S->sum), free(S->tmin), free(S->atten), free(S->emit), free(S->sumtot);
}
#define HREC(S, f, i) ((S)->hrec[(int64_t)(f) * (S)->n + (i)])
#define SREC(S, f, i) ((S)->srec[(int64_t)(f) * (S)->n + (i)])
#define ATT(S, c, d, i) ((S)->atten[((int64_t)(c) * (S)->depthCount + (d)) * (S)->n + (i)])
#define EMI(S, c, d, i) ((S)->emit[((int64_t)(c) * (S)->depthCount + (d)) * (S)->n + (i)])

typedef struct
{
  int64_t segments, draws, nan, zeroKilled;
  int64_t alive[64];
} LStats;

static inline float draw(uint32_t* seed, LStats* st)
{
  st->draws++;
  return orc_randf(seed);
}

/* Camera.cxx:894-953 + MapperPathTracer.cxx:283 */
static void st_create_ray(State* S, int64_t i, int64_t pixel, const cam_basis* cb, uint32_t* seed, LStats* st)
{
  st->draws += 2;
  v3 d = raygen(cb, pixel, seed);
  S->dx[i] = d.x, S->dy[i] = d.y, S->dz[i] = d.z;
  S->ox[i] = cb->pos.x, S->oy[i] = cb->pos.y, S->oz[i] = cb->pos.z;
  HREC(S, HR_T, i) = 0.f; /* Distance <- 0, Camera.cxx:899 */
  S->status[i] = 1u << BIT_SCT;
}
/* MapperPathTracer.cxx:287, :410-435 and SurfaceWorklets.h:98-111 */
static void st_intersect(State* S, int64_t i, int depth, const orc_scene* sc, int flags, LStats* st)
{
  S->sum[i] = 0.f;
  HREC(S, HR_T, i) = FLT_MAX; /* rays.Distance doubles as tmax and as hrec.T */
  S->tmin[i] = 0.001f;
  unsigned char sctr = S->status[i];
  if (sctr & (1u << BIT_SCT))
  {
    st->segments++;
    if (depth < 64)
      st->alive[depth]++;
    float rec[9];
    int hid[2];
    v3 o = V(S->ox[i], S->oy[i], S->oz[i]), d = V(S->dx[i], S->dy[i], S->dz[i]);
    int64_t prim = closest_hit(sc, o, d, S->tmin[i], FLT_MAX, flags, rec, hid);
    if (prim >= 0)
    {
      for (int f = 0; f < 9; f++)
        HREC(S, f, i) = rec[f];
      S->hid[i] = hid[0];
      S->hid[S->n + i] = hid[1];
      sctr |= (1u << BIT_HIT);
    }
  }
  /* CollectIntersecttWorklet */
  if (!((sctr & (1u << BIT_SCT)) && (sctr & (1u << BIT_HIT))))
  {
    sctr &= (unsigned char)~(1u << BIT_SCT);
    for (int c = 0; c < 3; c++)
    {
      ATT(S, c, depth, i) = 1.0f;
      EMI(S, c, depth, i) = 0.0f;
    }
  }
  sctr &= (unsigned char)~(1u << BIT_HIT);
  S->status[i] = sctr;
}
/* EmitWorklet.h:46-73, :112-135, :244-272 in launch order MapperPathTracer.cxx:470-477 */
static void st_materials(State* S, int64_t i, int depth, const orc_scene* sc, uint32_t* seed, LStats* st)
{
  unsigned char fin = S->status[i];
  if ((fin & (1u << BIT_FIN)) || !(fin & (1u << BIT_SCT)))
    return;
  int mt = sc->matType[S->hid[i]];
  int tt = sc->texType[S->hid[S->n + i]];
  v3 col = ld3(sc->tex + 3 * tt);
  v3 dir = V(S->dx[i], S->dy[i], S->dz[i]);
  v3 n = V(HREC(S, HR_NX, i), HREC(S, HR_NY, i), HREC(S, HR_NZ, i));
  if (mt == 0)
  { /* Lambertian */
    SREC(S, 6, i) = col.x, SREC(S, 7, i) = col.y, SREC(S, 8, i) = col.z;
    fin |= (1u << BIT_SCT);
    fin &= (unsigned char)~(1u << BIT_SPEC);
    for (int c = 0; c < 3; c++)
      EMI(S, c, depth, i) = 0.f;
  }
  else if (mt == 1)
  { /* DiffuseLight: fin &= (false << 3) -> 0 */
    v3 em = (vdot(n, dir) < 0.0f) ? col : V(0, 0, 0);
    fin = 0;
    EMI(S, 0, depth, i) = em.x, EMI(S, 1, depth, i) = em.y, EMI(S, 2, depth, i) = em.z;
  }
  else if (mt == 2)
  { /* Dielectric */
    float r = draw(seed, st);
    v3 nd = dielectric_scatter(dir, n, sc->refIdx, r);
    SREC(S, 6, i) = 1, SREC(S, 7, i) = 1, SREC(S, 8, i) = 1;
    SREC(S, 0, i) = HREC(S, HR_PX, i), SREC(S, 1, i) = HREC(S, HR_PY, i), SREC(S, 2, i) = HREC(S, HR_PZ, i);
    SREC(S, 3, i) = nd.x, SREC(S, 4, i) = nd.y, SREC(S, 5, i) = nd.z;
    fin |= (1u << BIT_SCT);
    fin |= (1u << BIT_SPEC);
    for (int c = 0; c < 3; c++)
      EMI(S, c, depth, i) = 0.f;
  }
  S->status[i] = fin;
}
/* PdfWorklet.h:19-21, :63-79, :112-137, :193-213 in launch order MapperPathTracer.cxx:488-502.
 * Runs for every pixel, alive or not. */
static void st_generate(State* S, int64_t i, const orc_scene* sc, uint32_t* seed, LStats* st)
{
  int which = (int)(draw(seed, st) * 3 + 1);
  if (which > 3)
    which = 3;
  S->which[i] = which;
  v3 p = V(HREC(S, HR_PX, i), HREC(S, HR_PY, i), HREC(S, HR_PZ, i));
  v3 g = V(S->gen[i], S->gen[S->n + i], S->gen[2 * S->n + i]);
  if (which <= 1)
  {
    float r1 = draw(seed, st);
    float r2 = draw(seed, st);
    v3 n = V(HREC(S, HR_NX, i), HREC(S, HR_NY, i), HREC(S, HR_NZ, i));
    onb_t uvw = onb_from_w(n);
    g = de_nan3(onb_local(&uvw, random_cosine_direction(r1, r2)));
  }
  if (which == 2)
  {
    for (int64_t l = 0; l < sc->nLightQuads; l++)
    {
      const int64_t* id = sc->lightQuadIds + 5 * l;
      v3 pt1 = ld3(sc->pts + 3 * id[1]);
      v3 pt2 = ld3(sc->pts + 3 * id[3]);
      float r1 = draw(seed, st);
      float r2 = draw(seed, st);
      float r3 = draw(seed, st);
      float y0 = pt1.y, y1 = pt1.y;
      v3 rp = V(pt1.x + r1 * (pt2.x - pt1.x), y0 + r2 * (y1 - y0), pt1.z + r3 * (pt2.z - pt1.z));
      g = vsub(rp, p);
    }
  }
  if (which == 3)
  {
    for (int64_t l = 0; l < sc->nLightSph; l++)
    {
      /* PdfWorklet.h:210: both draws are call arguments; GCC x86-64 evaluates right to left => r2 first */
      float r2 = draw(seed, st);
      float r1 = draw(seed, st);
      v3 center = ld3(sc->pts + 3 * sc->lightSphPt[l]);
      v3 direction = vsub(center, p);
      float d2 = vdot(direction, direction);
      onb_t uvw = onb_from_w(direction);
      g = de_nan3(onb_local(&uvw, random_to_sphere(sc->lightSphR[l], d2, r1, r2)));
    }
  }
  S->gen[i] = g.x, S->gen[S->n + i] = g.y, S->gen[2 * S->n + i] = g.z;
}
static inline float light_pdf_sum(const orc_scene* sc, v3 p, v3 g)
{
  float weight = (float)(1.0 / (double)(float)sc->lightables);
  float sum = 0.f;
  for (int64_t l = 0; l < sc->nLightQuads; l++)
  {
    const int64_t* id = sc->lightQuadIds + 5 * l;
    sum += weight * quad_pdf_value(p, g, ld3(sc->pts + 3 * id[1]), ld3(sc->pts + 3 * id[2]), ld3(sc->pts + 3 * id[3]),
                                   ld3(sc->pts + 3 * id[4]));
  }
  for (int64_t l = 0; l < sc->nLightSph; l++)
    sum += weight * sphere_pdf_value(p, g, ld3(sc->pts + 3 * sc->lightSphPt[l]), sc->lightSphR[l]);
  return sum;
}
/* PdfWorklet.h:274-316, :374-399, ScatterWorklet.h:67-117 in launch order MapperPathTracer.cxx:526-537 */
static void st_pdfs(State* S, int64_t i, int depth, const orc_scene* sc, uint32_t* seed, LStats* st)
{
  unsigned char fin = S->status[i];
  v3 p = V(HREC(S, HR_PX, i), HREC(S, HR_PY, i), HREC(S, HR_PZ, i));
  v3 g = V(S->gen[i], S->gen[S->n + i], S->gen[2 * S->n + i]);
  if (fin & (1u << BIT_SCT))
  {
    float sum = S->sum[i];
    float weight = (float)(1.0 / (double)(float)sc->lightables);
    for (int64_t l = 0; l < sc->nLightQuads; l++)
    {
      const int64_t* id = sc->lightQuadIds + 5 * l;
      sum += weight * quad_pdf_value(p, g, ld3(sc->pts + 3 * id[1]), ld3(sc->pts + 3 * id[2]),
                                     ld3(sc->pts + 3 * id[3]), ld3(sc->pts + 3 * id[4]));
    }
    (void)draw(seed, st); /* SpherePDFWorklet: int index = int(rand*list_size), unused */
    for (int64_t l = 0; l < sc->nLightSph; l++)
      sum += weight * sphere_pdf_value(p, g, ld3(sc->pts + 3 * sc->lightSphPt[l]), sc->lightSphR[l]);
    S->sum[i] = sum;
  }
  if (!(fin & (1u << BIT_FIN)))
  {
    v3 atten = V(1.0f, 1.0f, 1.0f);
    if (fin & (1u << BIT_SCT))
    {
      if (fin & (1u << BIT_SPEC))
      {
        atten = V(SREC(S, 6, i), SREC(S, 7, i), SREC(S, 8, i));
        S->ox[i] = SREC(S, 0, i), S->oy[i] = SREC(S, 1, i), S->oz[i] = SREC(S, 2, i);
        S->dx[i] = SREC(S, 3, i), S->dy[i] = SREC(S, 4, i), S->dz[i] = SREC(S, 5, i);
      }
      else
      {
        v3 n = V(HREC(S, HR_NX, i), HREC(S, HR_NY, i), HREC(S, HR_NZ, i));
        v3 albedo = V(SREC(S, 6, i), SREC(S, 7, i), SREC(S, 8, i));
        lambert_attenuation(n, g, S->sum[i], albedo, &atten);
        S->ox[i] = p.x, S->oy[i] = p.y, S->oz[i] = p.z;
        S->dx[i] = g.x, S->dy[i] = g.y, S->dz[i] = g.z;
      }
    }
    ATT(S, 0, depth, i) = atten.x, ATT(S, 1, depth, i) = atten.y, ATT(S, 2, depth, i) = atten.z;
  }
  fin &= (unsigned char)~(fin >> BIT_SCT);
  S->status[i] = fin;
}
/* MapperPathTracer.cxx:328-350 */
static void st_composite(State* S, int64_t i, float* canvas4, LStats* st)
{
  int D = S->depthCount;
  float L[3];
  for (int c = 0; c < 3; c++)
  {
    float l = EMI(S, c, D - 1, i) + 0.0f;
    for (int d = D - 2; d >= 0; d--)
    {
      l = ATT(S, c, d, i) * l;
      l = EMI(S, c, d, i) + l;
    }
    L[c] = l;
  }
  if (L[0] != L[0] || L[1] != L[1] || L[2] != L[2])
    st->nan++;
  for (int c = 0; c < 3; c++)
  {
    S->sumtot[(int64_t)c * S->n + i] = L[c];
    canvas4[c] += L[c];
  }
}

/* ------------------------------------------------------------------------------ forward form */
/* One path sample in forward (throughput) form.  Same stage order, same draws while alive.
 * burnDepths: if nonzero, a terminated path keeps consuming the draws the reference would. */
static inline void burn_depths(const orc_scene* sc, uint32_t* seed, int count, LStats* st)
{
  for (int k = 0; k < count; k++)
  {
    int which = (int)(draw(seed, st) * 3 + 1);
    if (which > 3)
      which = 3;
    /* the generators loop over every light, dead pixel or not (PdfWorklet.h:122, :203): cosine 2 draws, 3 per light
     * quad, 2 per light sphere */
    int nd = (which <= 1) ? 2 : (which == 2 ? 3 * (int)sc->nLightQuads : 2 * (int)sc->nLightSph);
    for (int j = 0; j < nd; j++)
      (void)draw(seed, st);
  }
}
#define ORC_LOG_STRIDE 20
static void forward_sample(const orc_scene* sc, const cam_basis* cb, int64_t pixel, uint32_t* seed, int maxDepth,
                           int burn, int flags, float* L3, LStats* st, float* log)
{
  st->draws += 2;
  v3 d = raygen(cb, pixel, seed);
  v3 o = cb->pos;
  v3 T = V(1.f, 1.f, 1.f);
  v3 L = V(0, 0, 0);
  int terminated = 0;
  for (int depth = 0; depth < maxDepth; depth++)
  {
    st->segments++;
    if (depth < 64)
      st->alive[depth]++;
    float rec[9];
    int hid[2];
    int64_t prim = closest_hit(sc, o, d, 0.001f, FLT_MAX, flags, rec, hid);
    float* lg = log ? log + (size_t)depth * ORC_LOG_STRIDE : NULL;
    if (lg)
    {
      memset(lg, 0, sizeof(float) * ORC_LOG_STRIDE);
      lg[0] = (float)prim, lg[1] = o.x, lg[2] = o.y, lg[3] = o.z, lg[4] = d.x, lg[5] = d.y, lg[6] = d.z;
      lg[7] = prim >= 0 ? rec[HR_T] : 0.f;
    }
    if (prim < 0)
    {
      L = vscale(T, 0.f);
      terminated = 1;
      if (burn)
        burn_depths(sc, seed, maxDepth - depth, st);
      break;
    }
    int mt = sc->matType[hid[0]];
    v3 col = ld3(sc->tex + 3 * sc->texType[hid[1]]);
    v3 n = V(rec[HR_NX], rec[HR_NY], rec[HR_NZ]);
    v3 p = V(rec[HR_PX], rec[HR_PY], rec[HR_PZ]);
    if (mt == 1)
    {
      v3 em = (vdot(n, d) < 0.0f) ? col : V(0, 0, 0);
      L = V(T.x * em.x, T.y * em.y, T.z * em.z);
      terminated = 1;
      if (burn)
        burn_depths(sc, seed, maxDepth - depth, st);
      break;
    }
    int specular = 0;
    v3 sdir = V(0, 0, 0);
    v3 albedo = col;
    if (mt == 2)
    {
      float r = draw(seed, st);
      sdir = dielectric_scatter(d, n, sc->refIdx, r);
      albedo = V(1, 1, 1);
      specular = 1;
    }
    else if (mt != 0)
    {
      /* unknown material type: no worklet touches the ray; it keeps its scatter bit and old srec.
       * Not reachable with the reference's tables; treat as lambertian with albedo col. */
    }
    /* generators */
    int which = (int)(draw(seed, st) * 3 + 1);
    if (which > 3)
      which = 3;
    v3 g = V(0, 0, 0);
    if (which <= 1)
    {
      float r1 = draw(seed, st);
      float r2 = draw(seed, st);
      onb_t uvw = onb_from_w(n);
      g = de_nan3(onb_local(&uvw, random_cosine_direction(r1, r2)));
    }
    else if (which == 2)
    {
      for (int64_t l = 0; l < sc->nLightQuads; l++)
      {
        const int64_t* id = sc->lightQuadIds + 5 * l;
        v3 pt1 = ld3(sc->pts + 3 * id[1]);
        v3 pt2 = ld3(sc->pts + 3 * id[3]);
        float r1 = draw(seed, st);
        float r2 = draw(seed, st);
        float r3 = draw(seed, st);
        float y0 = pt1.y, y1 = pt1.y;
        v3 rp = V(pt1.x + r1 * (pt2.x - pt1.x), y0 + r2 * (y1 - y0), pt1.z + r3 * (pt2.z - pt1.z));
        g = vsub(rp, p);
      }
    }
    else
    {
      for (int64_t l = 0; l < sc->nLightSph; l++)
      {
        float r2 = draw(seed, st);
        float r1 = draw(seed, st);
        v3 center = ld3(sc->pts + 3 * sc->lightSphPt[l]);
        v3 direction = vsub(center, p);
        float d2 = vdot(direction, direction);
        onb_t uvw = onb_from_w(direction);
        g = de_nan3(onb_local(&uvw, random_to_sphere(sc->lightSphR[l], d2, r1, r2)));
      }
    }
    /* pdfs */
    (void)draw(seed, st);
    if (specular)
    {
      /* atten = srec.A = 1 */
      o = p;
      d = sdir;
    }
    else
    {
      float sum = light_pdf_sum(sc, p, g);
      v3 atten;
      lambert_attenuation(n, g, sum, albedo, &atten);
      T = V(T.x * atten.x, T.y * atten.y, T.z * atten.z);
      if (lg)
      {
        lg[8] = (float)which, lg[9] = g.x, lg[10] = g.y, lg[11] = g.z, lg[12] = sum;
        lg[13] = vdot(n, unit_vector(g)), lg[14] = atten.x, lg[15] = T.x, lg[16] = n.x, lg[17] = n.y, lg[18] = n.z;
      }
      o = p;
      d = g;
      if ((flags & ORC_FLAG_KILL_ZERO_THROUGHPUT) && !burn && T.x == 0.f && T.y == 0.f && T.z == 0.f)
      {
        L = V(0, 0, 0);
        terminated = 1;
        st->zeroKilled++;
        break;
      }
    }
  }
  if (!terminated)
    L = vscale(T, 0.f); /* alive after maxDepth bounces: e[D-1] = 0 */
  if (L.x != L.x || L.y != L.y || L.z != L.z)
    st->nan++;
  st3(L3, L);
}

/* --------------------------------------------------------------------------------- entry points */
int orc_primary_hits(const orc_scene* sc, const orc_camera* cam, uint32_t seedOffset, int flags, int32_t* primId,
                     float* t)
{
  cam_basis cb = make_basis(cam);
  int64_t N = (int64_t)cam->W * cam->H;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N; i++)
  {
    uint32_t seed = (uint32_t)i + seedOffset;
    v3 d = raygen(&cb, i, &seed);
    float rec[9];
    int hid[2];
    rec[HR_T] = FLT_MAX;
    int64_t prim = closest_hit(sc, cb.pos, d, 0.001f, FLT_MAX, flags, rec, hid);
    primId[i] = (int32_t)prim;
    if (t)
      t[i] = prim >= 0 ? rec[HR_T] : FLT_MAX;
  }
  return 0;
}

/* ------------------------------------------------------------------------- direct G-buffers */
/* main.cc:402-422 (-direct): the MapperQuad family casts ONE un-jittered ray per pixel and shades the hit.
 *  * ray: Camera::PerspectiveRayGen (pathtracing/Camera.cxx:394-423, the reference's copy of VTK-m's generator):
 *    through the pixel's lower-left corner, dir = nlook + dx*((2i - W)/2) + dy*((2j - H)/2), zero components
 *    nudged, normalised by division;
 *  * hit: quads only (MapperQuad.cxx:103-113 extracts quads, no vertices/spheres), closest hit beyond t = 0, the
 *    normal flipped to oppose the ray (Surface.h:180-186, which restates VTK-m's quad intersector);
 *  * normals buffer: RayTracerNormals.cxx:137-140 -- (n.x, n.y, n.z, 1) where a quad is hit, untouched elsewhere;
 *  * albedo buffer: RayTracerAlbedo.cxx:100-143 -- with lightPosition = camera + 2*up (:154-155),
 *    L = normalize(lightPosition - p), V = normalize(camera - lookAt), cosTheta = clamp(n.L, 0, 1),
 *    R = normalize(2 (L.n) n - L), cosPhi = R.V: colour k = (cosPhi * R[k]) / (cosTheta * L[k]), alpha 1;
 *  * depth: the hit distance t along the unit ray (VTK-m's canvas stores a projected depth instead; that conversion
 *    lives inside VTK-m and is not restated), 0 where nothing is hit.
 * The colour image of the stock Phong shader ("direct-*.pnm") needs VTK-m's colour table and is out of scope. */
static inline v3 raygen_corner(const cam_basis* b, int64_t idx)
{
  int i = (int)((int32_t)idx % b->W);
  int j = (int)((int32_t)idx / b->W);
  v3 d = vadd(vadd(b->nlook, vscale(b->dx, (2.f * (float)i - (float)b->W) / 2.0f)),
              vscale(b->dy, (2.f * (float)j - (float)b->H) / 2.0f));
  if (d.x == 0.f)
    d.x += 0.0000001f;
  if (d.y == 0.f)
    d.y += 0.0000001f;
  if (d.z == 0.f)
    d.z += 0.0000001f;
  float m = sqrtf(vdot(d, d));
  return V(d.x / m, d.y / m, d.z / m);
}
void orc_direct_shade(const float* n3, const float* p3, const float* camPos3, const float* lookAt3, const float* upN3,
                      float* normals4, float* albedo4)
{
  const v3 n = ld3(n3), p = ld3(p3), cam = ld3(camPos3);
  const v3 lightPosition = vadd(cam, V(2.f * upN3[0], 2.f * upN3[1], 2.f * upN3[2]));
  const v3 L = vnormalize(vsub(lightPosition, p));
  const v3 Vd = vnormalize(vsub(cam, ld3(lookAt3)));
  float cosTheta = vdot(n, L);
  cosTheta = fminf(fmaxf(cosTheta, 0.f), 1.f);
  const float s = 2.f * vdot(L, n);
  const v3 R = vnormalize(vsub(vscale(n, s), L));
  const float cosPhi = vdot(R, Vd);
  if (normals4)
  {
    normals4[0] = n.x, normals4[1] = n.y, normals4[2] = n.z, normals4[3] = 1.f;
  }
  if (albedo4)
  {
    albedo4[0] = (cosPhi * R.x) / (cosTheta * L.x);
    albedo4[1] = (cosPhi * R.y) / (cosTheta * L.y);
    albedo4[2] = (cosPhi * R.z) / (cosTheta * L.z);
    albedo4[3] = 1.f;
  }
}
int orc_direct(const orc_scene* sc, const orc_camera* cam, float* normals4, float* albedo4, float* depth,
               int32_t* primId)
{
  if (cam->W <= 0 || cam->H <= 0)
    return -1;
  cam_basis cb = make_basis(cam);
  orc_scene quadsOnly = *sc;
  quadsOnly.nSph = 0;
  v3 up = ld3(cam->up);
  if (!(up.x == 0.f && up.y == 1.f && up.z == 0.f))
    up = vnormalize(up); /* Camera::SetUp, Camera.cxx:803-811 */
  const float upN[3] = { up.x, up.y, up.z };
  const int64_t N = (int64_t)cam->W * cam->H;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N; i++)
  {
    v3 d = raygen_corner(&cb, i);
    float rec[9];
    int hid[2];
    rec[HR_T] = FLT_MAX;
    int64_t prim = closest_hit(&quadsOnly, cb.pos, d, 0.f, FLT_MAX, 0, rec, hid);
    if (primId)
      primId[i] = (int32_t)prim;
    if (depth)
      depth[i] = prim >= 0 ? rec[HR_T] : 0.f;
    if (normals4)
      normals4[4 * i] = normals4[4 * i + 1] = normals4[4 * i + 2] = normals4[4 * i + 3] = 0.f;
    if (albedo4)
      albedo4[4 * i] = albedo4[4 * i + 1] = albedo4[4 * i + 2] = albedo4[4 * i + 3] = 0.f;
    if (prim >= 0)
      orc_direct_shade(rec + HR_NX, rec + HR_PX, cam->pos, cam->lookAt, upN, normals4 ? normals4 + 4 * i : NULL,
                       albedo4 ? albedo4 + 4 * i : NULL);
  }
  return 0;
}
void orc_raygen_corner(const orc_camera* cam, int64_t idx, float* dir3)
{
  cam_basis b = make_basis(cam);
  st3(dir3, raygen_corner(&b, idx));
}

static void merge_stats(orc_stats* dst, const LStats* s)
{
  dst->segments += s->segments;
  dst->rngDraws += s->draws;
  dst->nanSamples += s->nan;
  dst->zeroKilled += s->zeroKilled;
  for (int k = 0; k < 64; k++)
    dst->aliveAtDepth[k] += s->alive[k];
}

int orc_render(const orc_scene* sc, const orc_camera* cam, int spp, int sampleBegin, int maxDepth, uint32_t seedOffset,
               int mode, int flags, int nThreads, float* rgba, orc_stats* stats)
{
  if (spp < 0 || maxDepth < 1 || cam->W <= 0 || cam->H <= 0)
    return -1;
  if (mode != ORC_MODE_FORWARD_FAST && sampleBegin != 0)
    return -2;
  const int64_t N = (int64_t)cam->W * cam->H;
  cam_basis cb = make_basis(cam);
  orc_stats total;
  memset(&total, 0, sizeof(total));
  total.paths = N * (int64_t)spp;
  memset(rgba, 0, sizeof(float) * 4 * (size_t)N);
#ifdef _OPENMP
  omp_set_num_threads(nThreads > 0 ? nThreads : omp_get_num_procs());
#else
  (void)nThreads;
#endif

  if (mode == ORC_MODE_PASSES)
  {
    /* The reference's structure: one full-canvas loop per worklet, MapperPathTracer.cxx:278-350. */
    State S;
    if (state_alloc(&S, N, maxDepth))
      return -3;
    uint32_t* seeds = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)N);
    for (int64_t i = 0; i < N; i++)
      seeds[i] = (uint32_t)i + seedOffset; /* MapperPathTracer.cxx:265-267 */
#pragma omp parallel
    {
      LStats ls;
      memset(&ls, 0, sizeof(ls));
      for (int s = 0; s < spp; s++)
      {
#pragma omp for schedule(static)
        for (int64_t i = 0; i < N; i++)
          st_create_ray(&S, i, i, &cb, &seeds[i], &ls);
        for (int depth = 0; depth < maxDepth; depth++)
        {
#pragma omp for schedule(static)
          for (int64_t i = 0; i < N; i++)
            st_intersect(&S, i, depth, sc, flags, &ls);
#pragma omp for schedule(static)
          for (int64_t i = 0; i < N; i++)
            st_materials(&S, i, depth, sc, &seeds[i], &ls);
#pragma omp for schedule(static)
          for (int64_t i = 0; i < N; i++)
            st_generate(&S, i, sc, &seeds[i], &ls);
#pragma omp for schedule(static)
          for (int64_t i = 0; i < N; i++)
            st_pdfs(&S, i, depth, sc, &seeds[i], &ls);
        }
#pragma omp for schedule(static)
        for (int64_t i = 0; i < N; i++)
          st_composite(&S, i, rgba + 4 * i, &ls);
      }
#pragma omp critical
      merge_stats(&total, &ls);
    }
    free(seeds);
    state_free(&S);
  }
  else
  {
    int err = 0;
#pragma omp parallel
    {
      LStats ls;
      memset(&ls, 0, sizeof(ls));
      State S;
      int have_state = 0;
      if (mode == ORC_MODE_FUSED)
      {
        if (state_alloc(&S, 1, maxDepth))
          err = 1;
        else
          have_state = 1;
      }
#pragma omp for schedule(dynamic, 256)
      for (int64_t i = 0; i < N; i++)
      {
        float* px = rgba + 4 * i;
        if (mode == ORC_MODE_FUSED)
        {
          if (!have_state)
            continue;
          uint32_t seed = (uint32_t)i + seedOffset;
          memset(S.hrec, 0, sizeof(float) * 9);
          memset(S.gen, 0, sizeof(float) * 3);
          memset(S.srec, 0, sizeof(float) * 9);
          S.hid[0] = S.hid[1] = 0;
          for (int s = 0; s < spp; s++)
          {
            st_create_ray(&S, 0, i, &cb, &seed, &ls);
            for (int depth = 0; depth < maxDepth; depth++)
            {
              st_intersect(&S, 0, depth, sc, flags, &ls);
              st_materials(&S, 0, depth, sc, &seed, &ls);
              st_generate(&S, 0, sc, &seed, &ls);
              st_pdfs(&S, 0, depth, sc, &seed, &ls);
            }
            st_composite(&S, 0, px, &ls);
          }
        }
        else if (mode == ORC_MODE_FORWARD_BURN)
        {
          uint32_t seed = (uint32_t)i + seedOffset;
          for (int s = 0; s < spp; s++)
          {
            float L[3];
            forward_sample(sc, &cb, i, &seed, maxDepth, 1, flags, L, &ls, NULL);
            px[0] += L[0], px[1] += L[1], px[2] += L[2];
          }
        }
        else
        {
          for (int s = 0; s < spp; s++)
          {
            uint32_t seed = (uint32_t)i + seedOffset + (uint32_t)(sampleBegin + s) * ORC_GOLDEN;
            float L[3];
            forward_sample(sc, &cb, i, &seed, maxDepth, 0, flags, L, &ls, NULL);
            px[0] += L[0], px[1] += L[1], px[2] += L[2];
          }
        }
      }
      if (have_state)
        state_free(&S);
#pragma omp critical
      merge_stats(&total, &ls);
    }
    if (err)
      return -3;
  }
  if (stats)
    *stats = total;
  return 0;
}

/* Debug/test hook: one forward path sample from an explicit RNG state with a per-depth log
 * (ORC_LOG_STRIDE floats per depth: prim, o(3), d(3), t, which, g(3), lightPdfSum, cos, atten.x, T.x, n(3)). */
int orc_trace_path(const orc_scene* sc, const orc_camera* cam, int64_t pixel, uint32_t rngState, int maxDepth,
                   int flags, float* L3, float* log)
{
  cam_basis cb = make_basis(cam);
  LStats ls;
  memset(&ls, 0, sizeof(ls));
  uint32_t seed = rngState;
  forward_sample(sc, &cb, pixel, &seed, maxDepth, 0, flags, L3, &ls, log);
  return (int)ls.segments;
}

/* main.cc:253-287 */
void orc_normalize(const float* rgbaSum, int64_t n, int spp, float* out)
{
  float sc = (float)spp;
  for (int64_t i = 0; i < 4 * n; i++)
  {
    float v = rgbaSum[i];
    if ((i & 3) != 3 && !(v == v))
      v = 0;
    out[i] = sqrtf(v / sc);
  }
}
