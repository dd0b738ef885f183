/*
 * b2pt_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the reference's Monte-Carlo hot path
 * (m-kim/raytracingtherestofyourlife, MapperPathTracer::RenderCells and the
 * worklets it launches).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (libb2pt.so) never links, imports or calls it.
 *
 * PARITY STATUS: pinned to the reference's own code at worklet level; the residue is "parity unpinned".
 * The reference ships no tests, golden vectors or fixtures for this path and cannot be built as a whole here
 * (it needs VTK-m, which is absent).  But its hot-path arithmetic lives in header-only worklets;
 * oracle/ref_harness.cxx compiles those from /root/reference against a minimal VTK-m stand-in
 * (oracle/vtkm_min/) and runs them in the reference's launch order.  This oracle's PASSES mode is bit-identical
 * to that harness (images, segment counts, primary distances; tests/test_ref_harness.py and the committed
 * "refworklets" vectors in tests/golden/).  Camera ray generation is pinned the same way: the class
 * Camera::RayGen is lifted out of the reference's Camera.cxx at build time and compiled into the harness.
 * The scene too: the reference's CornellBox.cpp is compiled where it lies (oracle/ref_scene.cxx) and its points,
 * quad/sphere ids and material indices equal orc_cornell_scene's bit for bit.
 * Still unpinned, because they live inside VTK-m: its math (Cross/Normalize/Min/Max, restated in vtkm_min from its
 * documented semantics) and its LinearBVH tree shape (affects 1 of 65536 primary rays through the non-planar
 * quad's leaf box).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#ifndef B2PT_ORACLE_H
#define B2PT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Scene description: same arrays the reference builds in CornellBox.cpp:141-418
 * and extracts in MapperPathTracer.cxx:178-197 (QuadIds = Vec<Id,5>(cell,p0..p3)). */
typedef struct orc_scene
{
  int64_t nPts;
  const float* pts; /* 3*nPts */
  int64_t nQuads;
  const int64_t* quadIds; /* 5*nQuads */
  int64_t nSph;
  const int64_t* sphPt; /* nSph point ids */
  const float* sphR;    /* nSph radii */
  const int64_t* matIdxQ; /* nQuads */
  const int64_t* texIdxQ;
  const int64_t* matIdxS; /* nSph */
  const int64_t* texIdxS;
  int nMatType;
  const int* matType; /* 0 lambertian, 1 diffuse light, 2 dielectric */
  int nTexType;
  const int* texType;
  int nTex;
  const float* tex; /* 3*nTex */
  int64_t nLightQuads;
  const int64_t* lightQuadIds; /* 5*nLightQuads, MapperPathTracer.cxx:141-142 */
  int64_t nLightSph;
  const int64_t* lightSphPt; /* MapperPathTracer.cxx:145-146 */
  const float* lightSphR;    /* = sphR[i], PdfWorklet.h:205 */
  int lightables;            /* MapperPathTracer.cxx:218 (=2) */
  float refIdx;              /* MapperPathTracer.cxx:467 (=1.5) */
} orc_scene;

typedef struct orc_camera
{
  float pos[3], lookAt[3], up[3];
  float fovDeg;
  int W, H;
} orc_camera;

/* Execution modes (all render the same estimator):
 *  0 PASSES       stage-major loops over the whole canvas, one loop per worklet,
 *                 full depth, no early exit -- the structure of
 *                 MapperPathTracer.cxx:278-350.  This is the CPU baseline.
 *  1 FUSED        pixel-major, same stage functions, bitwise identical to 0.
 *  2 FORWARD_BURN forward throughput form, per-pixel persistent RNG stream; a dead
 *                 path burns the draws the reference would still consume, so
 *                 every trajectory equals modes 0/1 (radiance differs only by
 *                 product association, ~1e-7 relative).
 *  3 FORWARD_FAST forward form, one RNG stream per (pixel, sample):
 *                 state0 = pixel + seedOffset + sample*0x9E3779B9, no burn.  This
 *                 is the stream definition of the GPU production path.
 */
enum
{
  ORC_MODE_PASSES = 0,
  ORC_MODE_FUSED = 1,
  ORC_MODE_FORWARD_BURN = 2,
  ORC_MODE_FORWARD_FAST = 3
};

#define ORC_FLAG_KILL_ZERO_THROUGHPUT 1 /* only meaningful in mode 3 */
#define ORC_FLAG_NO_AABB_GATE 2 /* brute force: skip the per-primitive leaf-AABB slab test the reference's BVH implies */

typedef struct orc_stats
{
  int64_t paths;         /* path samples rendered */
  int64_t segments;      /* live ray segments traced (status bit 3 set when intersect runs) */
  int64_t rngDraws;      /* RNG draws consumed */
  int64_t nanSamples;    /* path samples whose radiance had a NaN channel */
  int64_t zeroKilled;    /* paths terminated by ORC_FLAG_KILL_ZERO_THROUGHPUT */
  int64_t aliveAtDepth[64]; /* live segments per depth (first 64 depths) */
} orc_stats;

/* wangXor.h:30-38, 55-59 */
uint32_t orc_wang32(uint32_t* state);
float orc_randf(uint32_t* state);
/* MapperPathTracer.cxx:60-75 */
uint32_t orc_wang_init(uint32_t x);

/* CornellBox.cpp:141-418: fills caller arrays; returns 0. Sizes: 89 points, 22 quads, 1 sphere. */
int orc_cornell_scene(float* pts /*3*89*/, int64_t* quadIds /*5*22*/, int64_t* sphPt /*1*/, float* sphR /*1*/,
                      int64_t* matIdxQ /*22*/, int64_t* texIdxQ /*22*/, int64_t* matIdxS /*1*/,
                      int64_t* texIdxS /*1*/, int* matType /*5*/, int* texType /*5*/, float* tex /*3*4*/);

/* Camera.cxx:438-476 (RayGen ctor) -> nlook, delta_x, delta_y (9 floats) */
void orc_camera_basis(const orc_camera* cam, float* nlook3, float* dx3, float* dy3);

/* Camera.cxx:483-524 for pixel idx with RNG state *seed (advanced by 2 draws) */
void orc_raygen(const orc_camera* cam, int64_t idx, uint32_t* seed, float* dir3);

/* Closest hit of one ray: BVHTraverser.h:128-227 semantics restated as brute force in
 * primitive index order (quads first, then spheres with the updated tmax).
 * Returns primitive id: quad q -> q, sphere s -> nQuads+s, miss -> -1. hrec9 = (u,v,t,nx,ny,nz,px,py,pz). */
int64_t orc_closest_hit(const orc_scene* sc, const float* o3, const float* d3, float tmin, float tmax, int flags,
                        float* hrec9, int* hid2);

/* Sample-0 primary rays with seeds[i] = i + seedOffset: hit primitive id and t per pixel. */
int orc_primary_hits(const orc_scene* sc, const orc_camera* cam, uint32_t seedOffset, int flags, int32_t* primId,
                     float* t);

/* -direct G-buffers (main.cc:402-422; raytracing/RayTracerNormals.cxx:47-143, RayTracerAlbedo.cxx:100-143): one
 * un-jittered ray per pixel (Camera::PerspectiveRayGen), closest QUAD hit, the two Shade rules; see b2pt_oracle.c.
 * normals4 / albedo4: W*H*4 floats, depth: W*H hit distances (0 = nothing hit), primId: W*H; any may be NULL. */
int orc_direct(const orc_scene* sc, const orc_camera* cam, float* normals4, float* albedo4, float* depth,
               int32_t* primId);
/* the per-pixel colour rules alone (for pinning against the reference's Shade worklets) and the pixel ray */
void orc_direct_shade(const float* n3, const float* p3, const float* camPos3, const float* lookAt3, const float* upN3,
                      float* normals4, float* albedo4);
void orc_raygen_corner(const orc_camera* cam, int64_t idx, float* dir3);

/* Render spp samples [sampleBegin, sampleBegin+spp) at maxDepth; rgba = un-normalised sum over samples
 * (MapperPathTracer.cxx:350), W*H*4 floats, alpha lane 0. In modes 0-2 sampleBegin must be 0. */
int orc_render(const orc_scene* sc, const orc_camera* cam, int spp, int sampleBegin, int maxDepth, uint32_t seedOffset,
               int mode, int flags, int nThreads, float* rgba, orc_stats* stats);

/* Debug/test hook: one forward path sample (mode-3 semantics) from an explicit RNG state; log gets 20 floats per
 * depth (prim, o, d, t, which, g, lightPdfSum, cos(n,g), atten.x, T.x, n). Returns segments traced. */
int orc_trace_path(const orc_scene* sc, const orc_camera* cam, int64_t pixel, uint32_t rngState, int maxDepth,
                   int flags, float* L3, float* log);

/* main.cc:253-287: out = sqrt(de_nan(sum)/spp) per channel (alpha passed through the same way). */
void orc_normalize(const float* rgbaSum, int64_t n, int spp, float* out);

/* Single light-sampling building blocks exposed for unit tests. */
float orc_quad_pdf_value(const float* o3, const float* v3, const float* q, const float* r, const float* s,
                         const float* t); /* PdfWorklet.h:230-248 */
float orc_sphere_pdf_value(const float* o3, const float* v3, const float* c3, float radius); /* PdfWorklet.h:333-347 */
int orc_quad_hit(const float* o3, const float* d3, const float* v00, const float* v10, const float* v11,
                 const float* v01, float* u, float* v, float* t); /* Surface.h:30-161 */

#ifdef __cplusplus
}
#endif
#endif
