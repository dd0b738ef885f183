// b2pt_lbvh.cu -- GPU LBVH builder (sm_100a) for scenes too large for the kernel-parameter path.
//
// Replaces, on the device, what the reference gets from VTK-m's LinearBVH (call sites QuadIntersector.cxx:134,
// SphereIntersector.cxx:75: FindQuadAABBs / FindSphereAABBs -> Morton codes -> radix tree): 30-bit Morton codes of the
// primitive AABB centroids, one cub radix sort, Karras' parallel radix-tree construction (one thread per internal
// node), bottom-up AABB fit with one atomic counter per node, and an emit pass that writes the traversal's 32-byte
// node layout directly (sibling pairs adjacent at 2i+2 / 2i+3, leaves of up to kLbvhLeaf primitives = every radix
// subtree that covers at most kLbvhLeaf sorted primitives, whose slots are contiguous by construction).
// The host binned-SAH builder (b2pt_bvh.h) stays the default: its trees trace faster; this one builds ~two orders of
// magnitude faster (B2PT_FLAG_GPU_LBVH).  Closest hits do not depend on the tree (tests compare with brute force).
#include <algorithm>
#include <cfloat>

#include <cub/device/device_radix_sort.cuh>

#include "b2pt_lbvh.h"

namespace b2pt
{
namespace
{

constexpr int kLbvhLeaf = 4;

// primitive boxes as the host builder pads them (b2pt_bvh.h quad_aabb / sphere_aabb; AABBSurface.h:66-77)
__device__ __forceinline__ void prim_box(const B2Quad* quads, const B2Sphere* sph, int enc, float* lo, float* hi)
{
  if (enc >= 0)
  {
    const B2Quad& Q = quads[enc];
    for (int c = 0; c < 3; ++c)
    {
      const float v0 = Q.v00[c], v2 = Q.v11[c], v1 = Q.v11[c] + Q.e21[c], v3 = Q.v11[c] + Q.e23[c];
      const float l = fminf(fminf(v0, v1), fminf(v2, v3)), h = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
      const float eps = fmaxf(1e-6f, 1.0e-4f * (h - l));
      lo[c] = l - eps;
      hi[c] = h + eps;
    }
  }
  else
  {
    const B2Sphere& S = sph[~enc];
    for (int c = 0; c < 3; ++c)
    {
      const float l = S.c[c] - S.r, h = S.c[c] + S.r;
      const float eps = fmaxf(1e-6f, 1.0e-4f * (h - l));
      lo[c] = l - eps;
      hi[c] = h + eps;
    }
  }
}

__device__ __forceinline__ uint32_t expand10(uint32_t v)
{ // 10 bits -> every third bit
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}

__global__ void __launch_bounds__(256)
  k_lbvh_morton(const B2Quad* quads, const B2Sphere* sph, const int32_t* treeQuads, int nq, int n, float3 org,
                float3 invExt, uint32_t* keys, int32_t* encs)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const int enc = i < nq ? treeQuads[i] : ~(i - nq);
  float lo[3], hi[3];
  prim_box(quads, sph, enc, lo, hi);
  const float cx = (0.5f * (lo[0] + hi[0]) - org.x) * invExt.x, cy = (0.5f * (lo[1] + hi[1]) - org.y) * invExt.y,
              cz = (0.5f * (lo[2] + hi[2]) - org.z) * invExt.z;
  const uint32_t x = (uint32_t)fminf(fmaxf(cx * 1024.f, 0.f), 1023.f), y = (uint32_t)fminf(fmaxf(cy * 1024.f, 0.f), 1023.f),
                 z = (uint32_t)fminf(fmaxf(cz * 1024.f, 0.f), 1023.f);
  keys[i] = (expand10(x) << 2) | (expand10(y) << 1) | expand10(z);
  encs[i] = enc;
}

// length of the common prefix of the (key, index) pairs i and j; -1 outside [0,n)
__device__ __forceinline__ int delta(const uint32_t* keys, int n, int i, int j)
{
  if (j < 0 || j >= n)
    return -1;
  const uint32_t a = keys[i], b = keys[j];
  if (a != b)
    return __clz(a ^ b);
  return 32 + __clz((uint32_t)i ^ (uint32_t)j);
}

struct RadixNode
{
  int left, right; // child: >=0 internal node, <0 leaf ~position
  int first, last; // sorted positions covered
  int parent;
};

// Karras 2012, "Maximizing parallelism in the construction of BVHs, octrees and k-d trees", algorithm of fig. 4
__global__ void __launch_bounds__(256) k_lbvh_radix(const uint32_t* keys, int n, RadixNode* nodes, int* leafParent)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1)
    return;
  const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = delta(keys, n, i, i - d);
  int lmax = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin)
    lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta(keys, n, i, i + (l + t) * d) > dmin)
      l += t;
  const int j = i + l * d;
  const int dnode = delta(keys, n, i, j);
  int s = 0, t = l;
  do
  {
    t = (t + 1) >> 1;
    if (delta(keys, n, i, i + (s + t) * d) > dnode)
      s += t;
  } while (t > 1);
  const int gamma = i + s * d + min(d, 0);
  const int lo = min(i, j), hi = max(i, j);
  RadixNode nd;
  nd.first = lo;
  nd.last = hi;
  nd.left = (lo == gamma) ? ~gamma : gamma;
  nd.right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
  nd.parent = nodes[i].parent; // written by the parent's thread below (or -1 for the root); keep it
  nodes[i].left = nd.left, nodes[i].right = nd.right, nodes[i].first = lo, nodes[i].last = hi;
  if (nd.left >= 0)
    nodes[nd.left].parent = i;
  else
    leafParent[~nd.left] = i;
  if (nd.right >= 0)
    nodes[nd.right].parent = i;
  else
    leafParent[~nd.right] = i;
}

// bottom-up box fit: the second thread to arrive at a node merges its children's boxes
__global__ void __launch_bounds__(256)
  k_lbvh_fit(const B2Quad* quads, const B2Sphere* sph, const int32_t* encs, int n, const RadixNode* nodes,
             const int* leafParent, float4* nlo, float4* nhi, float4* plo, float4* phi, int* arrived)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n)
    return;
  float lo[3], hi[3];
  prim_box(quads, sph, encs[p], lo, hi);
  plo[p] = make_float4(lo[0], lo[1], lo[2], 0.f);
  phi[p] = make_float4(hi[0], hi[1], hi[2], 0.f);
  if (n == 1)
    return;
  int node = leafParent[p];
  while (node >= 0)
  {
    __threadfence();
    if (atomicAdd(&arrived[node], 1) == 0)
      return; // the sibling subtree is not finished yet
    const RadixNode nd = nodes[node];
    const float4 al = nd.left >= 0 ? __ldcg(&nlo[nd.left]) : __ldcg(&plo[~nd.left]);
    const float4 ah = nd.left >= 0 ? __ldcg(&nhi[nd.left]) : __ldcg(&phi[~nd.left]);
    const float4 bl = nd.right >= 0 ? __ldcg(&nlo[nd.right]) : __ldcg(&plo[~nd.right]);
    const float4 bh = nd.right >= 0 ? __ldcg(&nhi[nd.right]) : __ldcg(&phi[~nd.right]);
    nlo[node] = make_float4(fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z), 0.f);
    nhi[node] = make_float4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), 0.f);
    node = nd.parent;
  }
}

__device__ __forceinline__ B2BvhNode child_record(int c, const RadixNode* nodes, const float4* nlo, const float4* nhi,
                                                  const float4* plo, const float4* phi)
{
  B2BvhNode r;
  float4 lo, hi;
  if (c < 0)
  { // single primitive
    lo = plo[~c], hi = phi[~c];
    r.left = ~c;
    r.count = 1;
  }
  else
  {
    lo = nlo[c], hi = nhi[c];
    const int size = nodes[c].last - nodes[c].first + 1;
    if (size <= kLbvhLeaf)
    { // radix subtree small enough: one leaf over its contiguous sorted range
      r.left = nodes[c].first;
      r.count = size;
    }
    else
    {
      r.left = 2 * c + 2;
      r.count = 0;
    }
  }
  r.bmin[0] = lo.x, r.bmin[1] = lo.y, r.bmin[2] = lo.z;
  r.bmax[0] = hi.x, r.bmax[1] = hi.y, r.bmax[2] = hi.z;
  return r;
}

__global__ void __launch_bounds__(256)
  k_lbvh_emit(int n, const RadixNode* nodes, const float4* nlo, const float4* nhi, const float4* plo, const float4* phi,
              const B2Sphere* sph, const int32_t* encs, B2BvhNode* out, float4* leafSph)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
  { // leaf-ordered sphere geometry
    const int enc = encs[i];
    leafSph[i] = enc < 0 ? make_float4(sph[~enc].c[0], sph[~enc].c[1], sph[~enc].c[2], sph[~enc].r)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (n == 1)
  {
    if (i == 0)
      out[0] = child_record(~0, nodes, nlo, nhi, plo, phi);
    return;
  }
  if (i >= n - 1)
    return;
  if (i == 0)
    out[0] = child_record(0, nodes, nlo, nhi, plo, phi); // root record: inner with children at 2,3 (or one leaf)
  if (nodes[i].last - nodes[i].first + 1 <= kLbvhLeaf)
    return; // collapsed into a leaf of its parent's record (or unused)
  out[2 * i + 2] = child_record(nodes[i].left, nodes, nlo, nhi, plo, phi);
  out[2 * i + 3] = child_record(nodes[i].right, nodes, nlo, nhi, plo, phi);
}

// ---- refit: the tree keeps its topology, the boxes follow the primitives' current geometry.  One thread per node; a
// reachable leaf recomputes its box from its primitives (the builders' padded boxes, prim_box) and climbs: the second
// thread to arrive at an inner node merges the boxes of its two (adjacent) children -- the bottom-up pass of the LBVH
// build, on the final 32-byte node layout of either builder.
__global__ void __launch_bounds__(256)
  k_bvh_refit(B2BvhNode* nodes, const int32_t* parent, int* arrived, int nNodes, const int32_t* slots,
              const B2Quad* quads, const B2Sphere* sph, float4* leafSph)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nNodes)
    return;
  if (parent[i] == -2) // not part of the tree (layout holes)
    return;
  B2BvhNode nd = nodes[i];
  if (nd.count <= 0)
    return; // inner nodes are filled by their second child to arrive
  float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
  for (int k = 0; k < nd.count; ++k)
  {
    const int enc = slots[nd.left + k];
    float a[3], b[3];
    prim_box(quads, sph, enc, a, b);
    for (int c = 0; c < 3; ++c)
      lo[c] = fminf(lo[c], a[c]), hi[c] = fmaxf(hi[c], b[c]);
    if (enc < 0)
      leafSph[nd.left + k] = make_float4(sph[~enc].c[0], sph[~enc].c[1], sph[~enc].c[2], sph[~enc].r);
  }
  for (int c = 0; c < 3; ++c)
    nodes[i].bmin[c] = lo[c], nodes[i].bmax[c] = hi[c];
  int node = parent[i];
  while (node >= 0)
  {
    __threadfence();
    if (atomicAdd(&arrived[node], 1) == 0)
      return; // the sibling subtree is not finished yet
    const int l = nodes[node].left;
    const float4* c4 = reinterpret_cast<const float4*>(nodes + l);
    const float4 al = __ldcg(c4), ah = __ldcg(c4 + 1), bl = __ldcg(c4 + 2), bh = __ldcg(c4 + 3);
    nodes[node].bmin[0] = fminf(al.x, bl.x), nodes[node].bmin[1] = fminf(al.y, bl.y), nodes[node].bmin[2] = fminf(al.z, bl.z);
    nodes[node].bmax[0] = fmaxf(ah.x, bh.x), nodes[node].bmax[1] = fmaxf(ah.y, bh.y), nodes[node].bmax[2] = fmaxf(ah.z, bh.z);
    node = parent[node];
  }
}

template <class T>
cudaError_t grow(T*& p, size_t& cap, size_t n)
{
  if (n <= cap)
    return cudaSuccess;
  if (p)
    cudaFree(p);
  p = nullptr;
  cap = 0;
  cudaError_t e = cudaMalloc(&p, n * sizeof(T));
  if (e == cudaSuccess)
    cap = n;
  return e;
}

} // namespace

cudaError_t build_lbvh_device(const B2Quad* dQuads, const B2Sphere* dSph, const int32_t* dTreeQuads, int nTreeQuads,
                              int nSph, const float sceneLo[3], const float sceneHi[3], B2BvhNode* dNodesOut,
                              int32_t* dSlotsOut, float4* dLeafSphOut, cudaStream_t stream)
{
  const int n = nTreeQuads + nSph;
  if (n <= 0)
    return cudaSuccess;
  cudaError_t e;
  uint32_t *keys = nullptr, *keysSorted = nullptr;
  int32_t* encs = nullptr;
  RadixNode* radix = nullptr;
  int *leafParent = nullptr, *arrived = nullptr;
  float4 *nlo = nullptr, *nhi = nullptr, *plo = nullptr, *phi = nullptr;
  void* temp = nullptr;
  size_t cap = 0, tempBytes = 0;
  auto cleanup = [&]() {
    for (void* p : { (void*)keys, (void*)keysSorted, (void*)encs, (void*)radix, (void*)leafParent, (void*)arrived,
                     (void*)nlo, (void*)nhi, (void*)plo, (void*)phi, temp })
      if (p)
        cudaFree(p);
  };
#define LB(x)                                                                                                          \
  if ((e = (x)) != cudaSuccess)                                                                                        \
  {                                                                                                                    \
    cleanup();                                                                                                         \
    return e;                                                                                                          \
  }
  cap = 0; LB(grow(keys, cap, (size_t)n));
  cap = 0; LB(grow(keysSorted, cap, (size_t)n));
  cap = 0; LB(grow(encs, cap, (size_t)n));
  cap = 0; LB(grow(radix, cap, (size_t)std::max(n - 1, 1)));
  cap = 0; LB(grow(leafParent, cap, (size_t)n));
  cap = 0; LB(grow(arrived, cap, (size_t)std::max(n - 1, 1)));
  cap = 0; LB(grow(nlo, cap, (size_t)std::max(n - 1, 1)));
  cap = 0; LB(grow(nhi, cap, (size_t)std::max(n - 1, 1)));
  cap = 0; LB(grow(plo, cap, (size_t)n));
  cap = 0; LB(grow(phi, cap, (size_t)n));
  const float3 org = make_float3(sceneLo[0], sceneLo[1], sceneLo[2]);
  auto inv = [](float lo, float hi) { return hi > lo ? 1.0f / (hi - lo) : 0.f; };
  const float3 invExt =
    make_float3(inv(sceneLo[0], sceneHi[0]), inv(sceneLo[1], sceneHi[1]), inv(sceneLo[2], sceneHi[2]));
  const int grid = (n + 255) / 256;
  k_lbvh_morton<<<grid, 256, 0, stream>>>(dQuads, dSph, dTreeQuads, nTreeQuads, n, org, invExt, keys, encs);
  // sort (key, enc) pairs; the sorted enc array IS the slot array of the tree
  LB(cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, keys, keysSorted, encs, dSlotsOut, n, 0, 30, stream));
  LB(cudaMalloc(&temp, std::max<size_t>(tempBytes, 16)));
  LB(cub::DeviceRadixSort::SortPairs(temp, tempBytes, keys, keysSorted, encs, dSlotsOut, n, 0, 30, stream));
  if (n > 1)
  {
    LB(cudaMemsetAsync(radix, 0xff, sizeof(RadixNode) * (size_t)(n - 1), stream)); // parent = -1 everywhere
    LB(cudaMemsetAsync(arrived, 0, sizeof(int) * (size_t)(n - 1), stream));
    k_lbvh_radix<<<(n - 1 + 255) / 256, 256, 0, stream>>>(keysSorted, n, radix, leafParent);
  }
  k_lbvh_fit<<<grid, 256, 0, stream>>>(dQuads, dSph, dSlotsOut, n, radix, leafParent, nlo, nhi, plo, phi, arrived);
  k_lbvh_emit<<<grid, 256, 0, stream>>>(n, radix, nlo, nhi, plo, phi, dSph, dSlotsOut, dNodesOut, dLeafSphOut);
  LB(cudaGetLastError());
  LB(cudaStreamSynchronize(stream));
#undef LB
  cleanup();
  return cudaSuccess;
}

cudaError_t refit_bvh_device(B2BvhNode* dNodes, const int32_t* dParent, int* dArrived, int nNodes,
                             const int32_t* dSlots, const B2Quad* dQuads, const B2Sphere* dSph, float4* dLeafSph,
                             cudaStream_t stream)
{
  if (nNodes <= 0)
    return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(dArrived, 0, sizeof(int) * (size_t)nNodes, stream);
  if (e != cudaSuccess)
    return e;
  k_bvh_refit<<<(nNodes + 255) / 256, 256, 0, stream>>>(dNodes, dParent, dArrived, nNodes, dSlots, dQuads, dSph,
                                                        dLeafSph);
  return cudaGetLastError();
}

} // namespace b2pt
