// b2pt_scene.cpp -- host-side scene builders exported through the C-ABI (include/b2pt.h).
//
// These restate INPUT DATA of the reference, not kernels: the Cornell box of CornellBox.cpp:141-418 and the
// synthetic many-sphere scene of BASELINE.json configs[3].  Host float arithmetic is
// compiled without FMA contraction so the arrays are bit-identical to what the reference's x86-64 build
// produces (and to the oracle's independent restatement; tests/test_scene.py compares the two).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/b2pt.h"

namespace
{

struct P3
{
  float x, y, z;
};

// Accumulates points / quads the way CornellBox::buildQuad does (CornellBox.cpp:37-61): every vertex is
// divided by 555.0 in Float64 and narrowed.
class CornellBuilder
{
public:
  std::vector<float> pts;
  std::vector<int64_t> quadIds, matIdx, texIdx;
  int64_t cells = 0;

  void point(P3 p)
  {
    pts.push_back(static_cast<float>(static_cast<double>(p.x) / 555.0));
    pts.push_back(static_cast<float>(static_cast<double>(p.y) / 555.0));
    pts.push_back(static_cast<float>(static_cast<double>(p.z) / 555.0));
  }
  void quad(const P3 (&v)[4], int m, int t)
  {
    const int64_t first = static_cast<int64_t>(pts.size() / 3);
    quadIds.push_back(cells++);
    for (int k = 0; k < 4; ++k)
    {
      point(v[k]);
      quadIds.push_back(first + k);
    }
    matIdx.push_back(m);
    texIdx.push_back(t);
  }
  // CornellBox::buildBox (CornellBox.cpp:63-139): z=near, z=far, y=far, then z=near and z=far again.
  void box(P3 lo, P3 hi, int m, int t)
  {
    const P3 zn[4] = { { lo.x, lo.y, lo.z }, { hi.x, lo.y, lo.z }, { hi.x, hi.y, lo.z }, { lo.x, hi.y, lo.z } };
    const P3 zf[4] = { { lo.x, lo.y, hi.z }, { hi.x, lo.y, hi.z }, { hi.x, hi.y, hi.z }, { lo.x, hi.y, hi.z } };
    const P3 top[4] = { { lo.x, hi.y, lo.z }, { hi.x, hi.y, lo.z }, { hi.x, hi.y, hi.z }, { lo.x, hi.y, hi.z } };
    quad(zn, m, t);
    quad(zf, m, t);
    quad(top, m, t);
    quad(zn, m, t);
    quad(zf, m, t);
  }
};

// CornellBox::invert (CornellBox.cpp:10-35): Translate(265,0,295) * Transpose(Rotate(-15 deg about +y)),
// vtkm::Transform3DRotate evaluated in Float32, matrix*vector rows as 4-term left-to-right dot products.
void tall_box_transform(P3 (&v)[4])
{
  const float rad = 0.01745329251994329547f * -15.f;
  const float s = std::sin(rad), c = std::cos(rad);
  const float omc = 1 - c;
  const float ax[3] = { 0.f, 1.f, 0.f };
  float rot[4][4] = {};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      rot[i][j] = ax[i] * ax[j] * omc;
  rot[0][0] += c, rot[1][1] += c, rot[2][2] += c;
  rot[0][1] -= ax[2] * s, rot[0][2] += ax[1] * s;
  rot[1][0] += ax[2] * s, rot[1][2] -= ax[0] * s;
  rot[2][0] -= ax[1] * s, rot[2][1] += ax[0] * s;
  rot[3][3] = 1.f;
  float tr[4][4] = {};
  tr[0][0] = tr[1][1] = tr[2][2] = tr[3][3] = 1.f;
  tr[0][3] = 265.f, tr[1][3] = 0.f, tr[2][3] = 295.f;
  float m[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
    {
      float acc = 0.f;
      for (int k = 0; k < 4; ++k)
        acc += tr[i][k] * rot[j][k]; // rot transposed
      m[i][j] = acc;
    }
  for (auto& p : v)
  {
    const float h[4] = { p.x, p.y, p.z, 1.f };
    float o[3];
    for (int r = 0; r < 3; ++r)
      o[r] = ((m[r][0] * h[0] + m[r][1] * h[1]) + m[r][2] * h[2]) + m[r][3] * h[3];
    p = { o[0], o[1], o[2] };
  }
}

inline uint32_t wang_step(uint32_t& s)
{
  s = (s ^ 61u) ^ (s >> 16);
  s *= 9u;
  s = s ^ (s >> 4);
  s *= 0x27d4eb2du;
  s = s ^ (s >> 15);
  return s;
}
inline float wang_float(uint32_t& s) { return static_cast<float>(wang_step(s)) / 4294967295.f; }

void material_tables(int* matType, int* texType, float* tex)
{
  // CornellBox.cpp:143-161
  const float colours[12] = { 0.65f, 0.05f, 0.05f, 0.73f, 0.73f, 0.73f, 0.12f, 0.45f, 0.15f, 15.f, 15.f, 15.f };
  const int mt[5] = { 0, 0, 0, 1, 2 };
  const int tt[5] = { 0, 1, 2, 3, 0 };
  std::memcpy(tex, colours, sizeof(colours));
  std::memcpy(matType, mt, sizeof(mt));
  std::memcpy(texType, tt, sizeof(tt));
}

} // namespace

extern "C" int b2pt_scene_cornell(float* pts, int64_t* quadIds, int64_t* spherePt, float* sphereR,
                                  int64_t* matIdxQuad, int64_t* texIdxQuad, int64_t* matIdxSph, int64_t* texIdxSph,
                                  int* matType, int* texType, float* tex)
{
  if (!pts || !quadIds || !spherePt || !sphereR || !matIdxQuad || !texIdxQuad || !matIdxSph || !texIdxSph ||
      !matType || !texType || !tex)
    return B2PT_ERR_BAD_VALUE;
  material_tables(matType, texType, tex);
  CornellBuilder b;
  {
    const P3 green[4] = { { 555, 0, 0 }, { 555, 555, 0 }, { 555, 555, 555 }, { 555, 0, 555 } };
    const P3 red[4] = { { 0, 0, 0 }, { 0, 555, 0 }, { 0, 555, 555 }, { 0, 0, 555 } };
    const P3 light[4] = { { 213, 554, 227 }, { 343, 554, 227 }, { 343, 554, 332 }, { 213, 554, 332 } };
    const P3 ceiling[4] = { { 0, 555, 0 }, { 555, 555, 0 }, { 555, 555, 555 }, { 0, 555, 555 } };
    const P3 floor_[4] = { { 0, 0, 0 }, { 555, 0, 0 }, { 555, 0, 555 }, { 0, 0, 555 } };
    const P3 back[4] = { { 0, 0, 555 }, { 555, 0, 555 }, { 555, 555, 555 }, { 0, 555, 555 } };
    b.quad(green, 2, 2);
    b.quad(red, 0, 0);
    b.quad(light, 3, 3);
    b.quad(ceiling, 1, 1);
    b.quad(floor_, 1, 1);
    b.quad(back, 1, 1);
  }
  {
    // tall box faces, CornellBox.cpp:264-353; the top face's first vertex (0,333,0) is the reference's typo
    const P3 faces[6][4] = {
      { { 0, 0, 165 }, { 165, 0, 165 }, { 165, 330, 165 }, { 0, 330, 165 } },
      { { 0, 0, 0 }, { 165, 0, 0 }, { 165, 330, 0 }, { 0, 330, 0 } },
      { { 165, 0, 0 }, { 165, 330, 0 }, { 165, 330, 165 }, { 165, 0, 165 } },
      { { 0, 0, 0 }, { 0, 330, 0 }, { 0, 330, 165 }, { 0, 0, 165 } },
      { { 0, 333, 0 }, { 165, 330, 0 }, { 165, 330, 165 }, { 0, 330, 165 } },
      { { 0, 0, 0 }, { 165, 0, 0 }, { 165, 0, 165 }, { 0, 0, 165 } },
    };
    for (const auto& f : faces)
    {
      P3 v[4] = { f[0], f[1], f[2], f[3] };
      tall_box_transform(v);
      b.quad(v, 1, 1);
    }
  }
  // sphere centre: the VERTEX cell (cell 12, point 48), CornellBox.cpp:357-365; radius MapperPathTracer.cxx:182
  spherePt[0] = static_cast<int64_t>(b.pts.size() / 3);
  b.point({ -335, 90, 290 });
  b.cells++;
  sphereR[0] = static_cast<float>(90 / 555.0);
  matIdxSph[0] = 4;
  texIdxSph[0] = 0;
  {
    const float rad = 90.f;
    const P3 centre = { 135, 90, 290 };
    b.box({ centre.x - rad, 0, centre.z - rad }, { centre.x + rad, 180, centre.z + rad }, 1, 1);
    b.box({ 50, 0, 50 }, { 450, 100, 100 }, 1, 1);
  }
  if (b.pts.size() != 3 * 89 || b.quadIds.size() != 5 * 22)
    return B2PT_ERR_STATE;
  std::memcpy(pts, b.pts.data(), b.pts.size() * sizeof(float));
  std::memcpy(quadIds, b.quadIds.data(), b.quadIds.size() * sizeof(int64_t));
  std::memcpy(matIdxQuad, b.matIdx.data(), b.matIdx.size() * sizeof(int64_t));
  std::memcpy(texIdxQuad, b.texIdx.data(), b.texIdx.size() * sizeof(int64_t));
  return B2PT_OK;
}

extern "C" int b2pt_scene_spheres(int64_t nSpheres, float* pts, int64_t* quadIds, int64_t* spherePt, float* sphereR,
                                  int64_t* matIdxQuad, int64_t* texIdxQuad, int64_t* matIdxSph, int64_t* texIdxSph,
                                  int* matType, int* texType, float* tex)
{
  if (nSpheres < 1 || !pts || !quadIds || !spherePt || !sphereR || !matIdxQuad || !texIdxQuad || !matIdxSph ||
      !texIdxSph || !matType || !texType || !tex)
    return B2PT_ERR_BAD_VALUE;
  material_tables(matType, texType, tex);
  for (int64_t k = 0; k < nSpheres; ++k)
  {
    uint32_t s = static_cast<uint32_t>(k);
    const float ra = wang_float(s), rb = wang_float(s), rc = wang_float(s), rd = wang_float(s);
    pts[3 * k + 0] = ra;
    pts[3 * k + 1] = 0.5f * rb;
    pts[3 * k + 2] = rc;
    sphereR[k] = 0.002f * (0.5f + rd);
    spherePt[k] = k;
    matIdxSph[k] = k % 3;
    texIdxSph[k] = k % 3;
  }
  const float q[8][3] = { { 0.3f, 0.98f, 0.3f }, { 0.7f, 0.98f, 0.3f }, { 0.7f, 0.98f, 0.7f }, { 0.3f, 0.98f, 0.7f },
                          { 0.f, 0.f, 0.f },     { 1.f, 0.f, 0.f },     { 1.f, 0.f, 1.f },     { 0.f, 0.f, 1.f } };
  for (int k = 0; k < 8; ++k)
    for (int c = 0; c < 3; ++c)
      pts[3 * (nSpheres + k) + c] = q[k][c];
  for (int f = 0; f < 2; ++f)
  {
    quadIds[5 * f] = f;
    for (int k = 0; k < 4; ++k)
      quadIds[5 * f + 1 + k] = nSpheres + 4 * f + k;
  }
  matIdxQuad[0] = 3, texIdxQuad[0] = 3; // emissive quad
  matIdxQuad[1] = 1, texIdxQuad[1] = 1; // white floor
  return B2PT_OK;
}
