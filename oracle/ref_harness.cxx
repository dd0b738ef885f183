// ref_harness.cxx -- TEST INFRASTRUCTURE (never linked, imported or called by the product).
//
// Drives the REFERENCE'S OWN header-only worklets, compiled from /root/reference where they lie, through the
// stage order of MapperPathTracer::RenderCellsImpl (MapperPathTracer.cxx:199-355).  The reference as a whole
// cannot be built here (it needs VTK-m, absent from this environment), but the arithmetic of the hot path lives
// in header-only functors that only need VTK-m's value types:
//     pathtracing/wangXor.h  vec3.h  onb.h  Record.h  AABBSurface.h  Surface.h  BVHTraverser.h
//     SurfaceWorklets.h  EmitWorklet.h  PdfWorklet.h  ScatterWorklet.h
// They are included below unmodified; oracle/vtkm_min/ supplies a minimal stand-in for the VTK-m headers they
// name (value types, Vec arithmetic, Min/Max/..., array portals; see vtkm_min/vtkm/Types.h for the semantics
// assumed).  What this file adds is only what VTK-m's dispatcher would do: call each worklet's operator() once
// per pixel with the arguments its ExecutionSignature lists, in the launch order of the reference.
//
// Camera ray generation: the class Camera::RayGen (Camera.cxx:425-524) lives in a .cxx file that needs all of VTK-m;
// the Makefile lifts exactly that class definition out of the reference file into _ref/camera_raygen_extract.inc
// (generated, git-ignored) and it is included below, so primary rays also come from the reference's own code.  The
// two statements that feed it (Look = normalize(LookAt - Position), Camera.cxx:911-912, and the constructor call,
// :935-940) are restated in b2ref_render.
//
// Not taken from the reference (restated, because it lives inside VTK-m):
//   * the LinearBVH build (VTK-m): a median-split tree over the reference's own leaf AABBs (AABBSurface.h), laid
//     out as BVHTraverser.h:45-69,182-221 consumes it.  The reference's traversal code itself runs on it.
//
// Built by oracle/Makefile target `ref` into oracle/_ref/libb2pt_refharness.so (git-ignored; travels to the GPU
// box); tests compare the C oracle (oracle/b2pt_oracle.c) against it bit for bit.
#include <algorithm>
#include <cfloat>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include <vtkm/Math.h>
#include <vtkm/VectorAnalysis.h>
#include <vtkm/cont/ArrayHandle.h>
#include <vtkm/worklet/WorkletMapField.h>

// ---- the reference's worklets, as they lie (include path: oracle/vtkm_min first, then /root/reference)
#include "pathtracing/wangXor.h"
#include "pathtracing/vec3.h"
#include "pathtracing/onb.h"
#include "pathtracing/Record.h"
#include "pathtracing/AABBSurface.h"
#include "pathtracing/Surface.h"
#include "pathtracing/BVHTraverser.h"
#include "pathtracing/SurfaceWorklets.h"
#include "pathtracing/EmitWorklet.h"
#include "pathtracing/PdfWorklet.h"
#include "pathtracing/ScatterWorklet.h"

#include "b2pt_oracle.h"

// ---- Camera::RayGen, lifted from the reference's Camera.cxx by the Makefile
namespace vtkm
{
namespace rendering
{
namespace pathtracing
{
class Camera
{
public:
  class RayGen;
};
#include "camera_raygen_extract.inc"
}
}
}

namespace
{
// Camera::SetUp (Camera.cxx:773-781) applied to the constructor's default Up = (0,1,0) (Camera.cxx:595-597): the
// vector is stored normalised unless it equals the default
inline vtkm::Vec<vtkm::Float32, 3> camera_up(const orc_camera* cam)
{
  vtkm::Vec<vtkm::Float32, 3> up(0.f, 1.f, 0.f);
  const vtkm::Vec<vtkm::Float32, 3> want(cam->up[0], cam->up[1], cam->up[2]);
  if (up[0] != want[0] || up[1] != want[1] || up[2] != want[2])
  {
    up = want;
    vtkm::Normalize(up);
  }
  return up;
}
using vtkm::Id;
using Vec4f = vtkm::Vec<vtkm::Float32, 4>;
using Id5 = vtkm::Vec<vtkm::Id, 5>;
using HitRecord = vtkm::Vec<vtkm::Float32, 9>;     // QuadIntersector.h HitRecord: composite of 9 Float32 arrays
using HitId = vtkm::Vec<vtkm::Int32, 2>;           // composite of matIdArray, texIdArray
using ScatterRecord = vtkm::Vec<vtkm::Float32, 9>; // composite of the specular_* buffers
using Serial = vtkm::cont::DeviceAdapterTagSerial;
template <typename T>
using Handle = vtkm::cont::ArrayHandle<T>;

struct Box
{
  float lo[3], hi[3];
};
// tree-shape variant (tests check that results do not depend on it): 0 median split on the widest axis,
// 1 the same with the children mirrored, 2 a left-deep chain in primitive order
int g_treeVariant = 0;

// Stand-in for vtkm::rendering::raytracing::LinearBVH::Construct: median split of the leaf AABBs' centroids
// along the widest axis, one primitive per leaf.  Layout: see vtkm_min/.../BoundingVolumeHierarchy.h.
struct FlatTree
{
  Handle<Vec4f> flat;
  Handle<Id> leafs;

  static float asFloat(vtkm::Int32 v)
  {
    float f;
    std::memcpy(&f, &v, 4);
    return f;
  }
  static Box merge(const std::vector<Box>& boxes, const std::vector<int>& ids, int b, int e)
  {
    Box r = boxes[ids[b]];
    for (int i = b + 1; i < e; ++i)
      for (int a = 0; a < 3; ++a)
      {
        r.lo[a] = std::min(r.lo[a], boxes[ids[i]].lo[a]);
        r.hi[a] = std::max(r.hi[a], boxes[ids[i]].hi[a]);
      }
    return r;
  }
  // returns the child reference (inner: first Vec4f index; leaf: -(offset)-1) and the subtree's box
  vtkm::Int32 build(const std::vector<Box>& boxes, std::vector<int>& ids, int b, int e, Box& out)
  {
    if (e - b == 1)
    {
      out = boxes[ids[b]];
      const Id off = leafs.GetNumberOfValues();
      leafs.Vector().push_back(1);
      leafs.Vector().push_back(ids[b]);
      return static_cast<vtkm::Int32>(-off - 1);
    }
    out = merge(boxes, ids, b, e);
    int axis = 0;
    float best = -1.f;
    for (int a = 0; a < 3; ++a)
    {
      float cmin = std::numeric_limits<float>::max(), cmax = -cmin;
      for (int i = b; i < e; ++i)
      {
        const float c = 0.5f * (boxes[ids[i]].lo[a] + boxes[ids[i]].hi[a]);
        cmin = std::min(cmin, c), cmax = std::max(cmax, c);
      }
      if (cmax - cmin > best)
        best = cmax - cmin, axis = a;
    }
    int m = (b + e) / 2;
    if (g_treeVariant == 2)
      m = e - 1;
    else
      std::stable_sort(ids.begin() + b, ids.begin() + e, [&](int x, int y) {
        const float cx = boxes[x].lo[axis] + boxes[x].hi[axis], cy = boxes[y].lo[axis] + boxes[y].hi[axis];
        return g_treeVariant == 1 ? cx > cy : cx < cy;
      });
    const Id node = flat.GetNumberOfValues();
    flat.Vector().resize(static_cast<size_t>(node + 4));
    Box lb, rb;
    const vtkm::Int32 l = build(boxes, ids, b, m, lb);
    const vtkm::Int32 r = build(boxes, ids, m, e, rb);
    std::vector<Vec4f>& F = flat.Vector();
    F[node + 0] = Vec4f(lb.lo[0], lb.lo[1], lb.lo[2], lb.hi[0]);
    F[node + 1] = Vec4f(lb.hi[1], lb.hi[2], rb.lo[0], rb.lo[1]);
    F[node + 2] = Vec4f(rb.lo[2], rb.hi[0], rb.hi[1], rb.hi[2]);
    F[node + 3] = Vec4f(asFloat(l), asFloat(r), 0.f, 0.f);
    return static_cast<vtkm::Int32>(node);
  }
  void construct(const std::vector<Box>& boxes)
  {
    std::vector<int> ids(boxes.size());
    for (size_t i = 0; i < ids.size(); ++i)
      ids[i] = static_cast<int>(i);
    if (boxes.size() == 1)
    { // single primitive: root with the primitive on the left and an empty leaf (count 0) on the right
      flat.Vector().resize(4);
      leafs.Vector() = { 1, 0, 0 };
      const Box& b = boxes[0];
      const float inf = std::numeric_limits<float>::infinity();
      flat.Vector()[0] = Vec4f(b.lo[0], b.lo[1], b.lo[2], b.hi[0]);
      flat.Vector()[1] = Vec4f(b.hi[1], b.hi[2], inf, inf);
      flat.Vector()[2] = Vec4f(inf, -inf, -inf, -inf);
      flat.Vector()[3] = Vec4f(asFloat(-1), asFloat(-3), 0.f, 0.f);
      return;
    }
    Box root;
    build(boxes, ids, 0, static_cast<int>(boxes.size()), root);
  }
};

template <typename T>
Handle<T> make_handle(const T* p, Id n)
{
  Handle<T> h;
  h.Vector().assign(p, p + n);
  return h;
}
} // namespace

extern "C" {

void b2ref_set_tree_variant(int v) { g_treeVariant = v; }

// wangXor.h:30-38, 55-59 through the reference's own functions
uint32_t b2ref_wang32(uint32_t* state) { return xorshiftWang::getWang32(*state); }
float b2ref_randf(uint32_t* state) { return xorshiftWang::getRandF(*state); }

// Camera::RayGen (Camera.cxx:425-524) for one pixel: seed is advanced by the two jitter draws
void b2ref_raygen(const orc_camera* cam, int64_t idx, uint32_t* seed, float* dir3)
{
  vec3 look = vec3(cam->lookAt[0], cam->lookAt[1], cam->lookAt[2]) - vec3(cam->pos[0], cam->pos[1], cam->pos[2]);
  vtkm::Normalize(look);
  const vtkm::rendering::pathtracing::Camera::RayGen raygen(cam->W, cam->H, cam->fovDeg, cam->fovDeg, look,
                                                            camera_up(cam), 0, cam->W, 0, 0);
  float dx = 0.f, dy = 0.f, dz = 0.f;
  vtkm::Id pixelIndex = 0;
  unsigned int s = *seed;
  raygen(idx, dx, dy, dz, s, pixelIndex);
  *seed = s;
  dir3[0] = dx, dir3[1] = dy, dir3[2] = dz;
}

// Surface.h:30-161 (QuadLeafIntersector::hit) on explicit vertices
int b2ref_quad_hit(const float* o, const float* d, const float* v00, const float* v10, const float* v11,
                   const float* v01, float* u, float* v, float* t)
{
  QuadLeafIntersector<Serial> q;
  return q.hit(vec3(o[0], o[1], o[2]), vec3(d[0], d[1], d[2]), vec3(v00[0], v00[1], v00[2]),
               vec3(v10[0], v10[1], v10[2]), vec3(v11[0], v11[1], v11[2]), vec3(v01[0], v01[1], v01[2]), *u, *v, *t)
    ? 1
    : 0;
}

// Surface.h:319-367 (SphereLeafIntersector::hit): rec9 = (u,v,t,nx,ny,nz,px,py,pz)
int b2ref_sphere_hit(const float* o, const float* d, float tmin, float tmax, const float* c, float radius,
                     float* rec9)
{
  SphereLeafIntersector<Serial> s;
  HitRecord rec(0.f);
  HitId hid(0);
  const bool h = s.hit(vec3(o[0], o[1], o[2]), vec3(d[0], d[1], d[2]), rec, hid, tmin, tmax, vec3(c[0], c[1], c[2]),
                       radius, 0, 0);
  for (int k = 0; k < 9; ++k)
    rec9[k] = rec[k];
  return h ? 1 : 0;
}

// The whole depth loop of MapperPathTracer.cxx:276-351 with the reference's worklets.  Every per-pixel launch is an
// OpenMP parallel-for (the stand-in for VTK-m's OpenMP device adapter; OMP_NUM_THREADS=1 = its Serial adapter).
// rgba: un-normalised sum over samples (alpha lane 0); segments: live rays entering `intersect`, summed.
// primId0/t0 (optional): per pixel, sample 0 / depth 0: hit material id pair is not a primitive id, so the
// harness reports the closest distance (Distance array after intersect) and whether the pixel stayed alive.
int b2ref_render(const orc_scene* sc, const orc_camera* cam, int spp, int maxDepth, float* rgba, int64_t* segments,
                 float* t0, uint8_t* hit0)
{
  const Id N = static_cast<Id>(cam->W) * cam->H;
  const int D = maxDepth;
  // ---- scene arrays (CornellBox.h members / MapperPathTracer::extract)
  Handle<vec3> coords;
  coords.Allocate(sc->nPts);
  for (Id i = 0; i < sc->nPts; ++i)
    coords.Vector()[i] = vec3(sc->pts[3 * i], sc->pts[3 * i + 1], sc->pts[3 * i + 2]);
  Handle<Id5> QuadIds;
  QuadIds.Allocate(sc->nQuads);
  for (Id q = 0; q < sc->nQuads; ++q)
    QuadIds.Vector()[q] = Id5(sc->quadIds[5 * q], sc->quadIds[5 * q + 1], sc->quadIds[5 * q + 2],
                              sc->quadIds[5 * q + 3], sc->quadIds[5 * q + 4]);
  Handle<Id> SphereIds = make_handle<Id>(reinterpret_cast<const Id*>(sc->sphPt), sc->nSph);
  Handle<vtkm::Float32> SphereRadii = make_handle<vtkm::Float32>(sc->sphR, sc->nSph);
  Handle<Id> MatIdx[2] = { make_handle<Id>(reinterpret_cast<const Id*>(sc->matIdxQ), sc->nQuads),
                           make_handle<Id>(reinterpret_cast<const Id*>(sc->matIdxS), sc->nSph) };
  Handle<Id> TexIdx[2] = { make_handle<Id>(reinterpret_cast<const Id*>(sc->texIdxQ), sc->nQuads),
                           make_handle<Id>(reinterpret_cast<const Id*>(sc->texIdxS), sc->nSph) };
  Handle<int> MatType = make_handle<int>(sc->matType, sc->nMatType);
  Handle<int> TexType = make_handle<int>(sc->texType, sc->nTexType);
  Handle<vec3> Tex;
  Tex.Allocate(sc->nTex);
  for (int i = 0; i < sc->nTex; ++i)
    Tex.Vector()[i] = vec3(sc->tex[3 * i], sc->tex[3 * i + 1], sc->tex[3 * i + 2]);
  // light lists (MapperPathTracer.cxx:141-148): point-id lists plus index lists [0..n)
  Handle<Id5> light_box_pointids;
  light_box_pointids.Allocate(sc->nLightQuads);
  Handle<Id> light_box_indices, light_sphere_indices;
  for (Id l = 0; l < sc->nLightQuads; ++l)
  {
    const int64_t* p = sc->lightQuadIds + 5 * l;
    light_box_pointids.Vector()[l] = Id5(p[0], p[1], p[2], p[3], p[4]);
    light_box_indices.Vector().push_back(l);
  }
  Handle<Id> light_sphere_pointids = make_handle<Id>(reinterpret_cast<const Id*>(sc->lightSphPt), sc->nLightSph);
  Handle<vtkm::Float32> light_sphere_radii = make_handle<vtkm::Float32>(sc->lightSphR, sc->nLightSph);
  for (Id l = 0; l < sc->nLightSph; ++l)
    light_sphere_indices.Vector().push_back(l);
  const int lightables = sc->lightables;

  // ---- buildBVH (MapperPathTracer.cxx:437-451): leaf AABBs by the reference's worklets, stand-in tree
  FlatTree quadTree, sphereTree;
  {
    std::vector<Box> boxes(static_cast<size_t>(sc->nQuads));
    ::detail::FindQuadAABBs fq;
    for (Id q = 0; q < sc->nQuads; ++q)
    {
      Box& b = boxes[q];
      fq(QuadIds.Vector()[q], b.lo[0], b.lo[1], b.lo[2], b.hi[0], b.hi[1], b.hi[2], coords.Portal());
    }
    if (sc->nQuads > 0)
      quadTree.construct(boxes);
    boxes.assign(static_cast<size_t>(sc->nSph), Box());
    ::detail::FindSphereAABBs fs;
    for (Id s = 0; s < sc->nSph; ++s)
    {
      Box& b = boxes[s];
      fs(SphereIds.Vector()[s], SphereRadii.Vector()[s], b.lo[0], b.lo[1], b.lo[2], b.hi[0], b.hi[1], b.hi[2],
         coords.Portal());
    }
    if (sc->nSph > 0)
      sphereTree.construct(boxes);
  }
  QuadExecWrapper quadWrap(QuadIds, MatIdx[0], TexIdx[0]);
  SphereExecWrapper sphereWrap(SphereIds, SphereRadii, MatIdx[1], TexIdx[1]);
  auto quadLeaf = quadWrap.PrepareForExecution(Serial());
  auto sphereLeaf = sphereWrap.PrepareForExecution(Serial());

  // ---- per-pixel state (Ray<Float32> arrays and the ChannelBuffers of MapperPathTracer.cxx:111-139)
  std::vector<vec3> origin(N, vec3(0.f)), dir(N, vec3(0.f)), generated(N, vec3(0.f));
  std::vector<HitRecord> hrec(N, HitRecord(0.f)); // component T aliases rays.Distance
  std::vector<HitId> hid(N, HitId(0));
  std::vector<ScatterRecord> srec(N, ScatterRecord(0.f));
  std::vector<vtkm::UInt8> status(N, 0);
  std::vector<int> which(N, 0);
  std::vector<float> sum_values(N, 0.f), tmin(N, 0.f);
  std::vector<unsigned int> seeds(N);
  Handle<vec3> attenuation, emitted;
  attenuation.Allocate(N * D);
  emitted.Allocate(N * D);
  std::vector<vec3> sumtotl(N, vec3(0.f));
  for (Id i = 0; i < N; ++i)
    seeds[i] = static_cast<unsigned int>(i); // MapperPathTracer.cxx:265-267 (CopyIf with a constant-true predicate)
  for (Id i = 0; i < 4 * N; ++i)
    rgba[i] = 0.f;
  int64_t segs = 0;
  const float HRT = static_cast<Id>(HR::T);
  (void)HRT;
  vtkm::rendering::pathtracing::BVHTraverser::Intersector traverse;

  // Camera.cxx:911-912 and :935-940: Look = normalize(LookAt - Position); RayGen(W, H, fov, fov, Look, Up, 0, W, 0, 0)
  vec3 look = vec3(cam->lookAt[0], cam->lookAt[1], cam->lookAt[2]) - vec3(cam->pos[0], cam->pos[1], cam->pos[2]);
  vtkm::Normalize(look);
  const vtkm::rendering::pathtracing::Camera::RayGen raygen(cam->W, cam->H, cam->fovDeg, cam->fovDeg, look,
                                                            camera_up(cam), 0, cam->W, 0, 0);
  for (int s = 0; s < spp; ++s)
  {
    // rayCam.CreateRays (Camera.cxx:880-960) + Status = 1<<3 (MapperPathTracer.cxx:283)
    _Pragma("omp parallel for schedule(static)")
    for (Id i = 0; i < N; ++i)
    {
      float dx = 0.f, dy = 0.f, dz = 0.f;
      vtkm::Id pixelIndex = 0;
      raygen(i, dx, dy, dz, seeds[i], pixelIndex);
      dir[i] = vec3(dx, dy, dz);
      origin[i] = vec3(cam->pos[0], cam->pos[1], cam->pos[2]);
      hrec[i][static_cast<Id>(HR::T)] = 0.f; // Distance <- 0 (Camera.cxx:899)
      status[i] = static_cast<vtkm::UInt8>(1UL << 3);
    }
    for (int depth = 0; depth < D; ++depth)
    {
      // MapperPathTracer.cxx:287 and ::intersect (:410-435)
      _Pragma("omp parallel for schedule(static) reduction(+ : segs)")
      for (Id i = 0; i < N; ++i)
      {
        sum_values[i] = 0.f;
        hrec[i][static_cast<Id>(HR::T)] = std::numeric_limits<float>::max(); // rays.Distance
        tmin[i] = static_cast<float>(0.001);
        if (status[i] & (1UL << 3))
          ++segs;
      }
      for (int pass = 0; pass < 2; ++pass) // quadIntersector.IntersectRays, then sphereIntersector.IntersectRays
      {
        const FlatTree& tree = pass == 0 ? quadTree : sphereTree;
        if (tree.flat.GetNumberOfValues() == 0)
          continue;
        _Pragma("omp parallel for schedule(static)")
        for (Id i = 0; i < N; ++i)
        {
          // FieldInOut arguments are loaded into locals, passed by reference, and stored back in order:
          // hrecs (whose T component is rays.Distance) first, rays.Distance (tmax) later.
          HitRecord h = hrec[i];
          float tmax = hrec[i][static_cast<Id>(HR::T)];
          if (pass == 0)
            traverse(i, origin[i], dir[i], h, hid[i], tmin[i], tmax, status[i], coords.Portal(), quadLeaf,
                     tree.flat.Portal(), tree.leafs.Portal());
          else
            traverse(i, origin[i], dir[i], h, hid[i], tmin[i], tmax, status[i], coords.Portal(), sphereLeaf,
                     tree.flat.Portal(), tree.leafs.Portal());
          hrec[i] = h;
          hrec[i][static_cast<Id>(HR::T)] = tmax;
        }
      }
      if (s == 0 && depth == 0 && t0 && hit0)
        _Pragma("omp parallel for schedule(static)")
        for (Id i = 0; i < N; ++i)
        {
          t0[i] = hrec[i][static_cast<Id>(HR::T)];
          hit0[i] = (status[i] & (1UL << 2)) ? 1 : 0;
        }
      {
        CollectIntersecttWorklet collect(N, depth);
        auto ep = emitted.Portal(), ap = attenuation.Portal();
        _Pragma("omp parallel for schedule(static)")
        for (Id i = 0; i < N; ++i)
          collect(i, status[i], ep, ap);
      }
      // ::applyMaterials (:453-479)
      {
        LambertianWorklet lmb(N, depth);
        DiffuseLightWorklet dl(N, depth);
        DielectricWorklet de(N, depth, 1.5, static_cast<vtkm::UInt32>(N));
        (void)sc->refIdx; // the reference hard-codes 1.5 (MapperPathTracer.cxx:467)
        _Pragma("omp parallel for schedule(static)")
        for (Id i = 0; i < N; ++i)
          lmb(i, origin[i], dir[i], hrec[i], hid[i], srec[i], status[i], Tex.Portal(), MatType.Portal(),
              TexType.Portal(), emitted.Portal());
        _Pragma("omp parallel for schedule(static)")
        for (Id i = 0; i < N; ++i)
          dl(i, origin[i], dir[i], hrec[i], hid[i], srec[i], status[i], Tex.Portal(), MatType.Portal(),
             TexType.Portal(), emitted.Portal());
        _Pragma("omp parallel for schedule(static)")
        for (Id i = 0; i < N; ++i)
          de(i, seeds[i], origin[i], dir[i], hrec[i], hid[i], srec[i], status[i], Tex.Portal(), MatType.Portal(),
             TexType.Portal(), emitted.Portal());
      }
      // ::generateRays (:481-503)
      {
        WorketletGenerateDir genDir(3); // WhichGenerateDir.cxx:10
        _Pragma("omp parallel for schedule(static)")
        for (Id i = 0; i < N; ++i)
          genDir(seeds[i], which[i]);
        CosineWorketletGenerateDir cosGen(1); // CosineGenerateDir.h:18
        _Pragma("omp parallel for schedule(static)")
        for (Id i = 0; i < N; ++i)
          cosGen(which[i], hrec[i], generated[i], seeds[i]);
        QuadWorkletGenerateDir quadGen(2); // QuadGenerateDir.h:22
        _Pragma("omp parallel for schedule(static)")
        for (Id i = 0; i < N; ++i)
          quadGen(which[i], hrec[i], generated[i], seeds[i], light_box_pointids.Portal(), light_box_indices.Portal(),
                  coords.Portal());
        SphereWorkletGenerateDir sphGen(3); // SphereGenerateDir.h:24
        _Pragma("omp parallel for schedule(static)")
        for (Id i = 0; i < N; ++i)
          sphGen(i, which[i], hrec[i], generated[i], seeds[i], light_sphere_pointids.Portal(),
                 light_sphere_indices.Portal(), coords.Portal(), light_sphere_radii.Portal());
      }
      // ::applyPDFs (:507-538)
      {
        QuadPDFWorklet quadPdf(lightables);
        _Pragma("omp parallel for schedule(static)")
        for (Id i = 0; i < N; ++i)
          quadPdf(i, origin[i], dir[i], hrec[i], status[i], sum_values[i], generated[i], seeds[i], quadLeaf,
                  light_box_pointids.Portal(), light_box_indices.Portal(), coords.Portal());
        SpherePDFWorklet sphPdf(lightables);
        _Pragma("omp parallel for schedule(static)")
        for (Id i = 0; i < N; ++i)
          sphPdf(i, origin[i], dir[i], hrec[i], status[i], sum_values[i], generated[i], seeds[i], sphereLeaf,
                 light_sphere_pointids.Portal(), light_sphere_indices.Portal(), coords.Portal(),
                 light_sphere_radii.Portal());
        PDFCosineWorklet pdfW(static_cast<int>(N), depth, static_cast<vtkm::UInt32>(N), lightables);
        _Pragma("omp parallel for schedule(static)")
        for (Id i = 0; i < N; ++i)
        {
          // rays.Origin / rays.Dir are passed twice (in: _1,_2; out: _8,_9); the later store wins
          vec3 ro = origin[i], rd = dir[i], oo = origin[i], od = dir[i];
          pdfW(i, ro, rd, hrec[i], srec[i], status[i], sum_values[i], generated[i], oo, od, attenuation.Portal());
          origin[i] = oo;
          dir[i] = od;
        }
      }
    }
    // compositing (MapperPathTracer.cxx:328-350): sumtotl = e[D-1] + 0; then a[d]*sumtotl, e[d]+sumtotl; cols += sumtotl
    const std::vector<vec3>& E = emitted.Vector();
    const std::vector<vec3>& A = attenuation.Vector();
    _Pragma("omp parallel for schedule(static)")
    for (Id i = 0; i < N; ++i)
    {
      sumtotl[i] = E[static_cast<size_t>((D - 1) * N + i)] + vec3(0.0f);
      for (int depth = D - 2; depth >= 0; --depth)
      {
        sumtotl[i] = A[static_cast<size_t>(depth * N + i)] * sumtotl[i];
        sumtotl[i] = E[static_cast<size_t>(depth * N + i)] + sumtotl[i];
      }
      for (int c = 0; c < 3; ++c)
        rgba[4 * i + c] = rgba[4 * i + c] + sumtotl[i][c];
    }
  }
  if (segments)
    *segments = segs;
  return 0;
}
} // extern "C"
