// b2pt_bvh.h -- host-side BVH builder (binned SAH) producing the compact 32-byte node layout the traversal
// kernel consumes (B2BvhNode: two 16-byte vector loads per node, sibling pairs adjacent = one 64-byte fetch).
//
// Replaces, for scenes too large for the kernel-parameter path, the reference's two per-shape VTK-m LinearBVH
// builds (pathtracing/QuadIntersector.cxx:134, pathtracing/SphereIntersector.cxx:75) with ONE tree over quads
// and spheres.  Primitive AABBs follow pathtracing/AABBSurface.h (quads padded by max(1e-6, 1e-4*extent));
// sphere boxes get the same relative padding (the reference leaves them unpadded) so that box culling is
// conservative with respect to the analytic sphere test under float rounding.
// The closest hit does not depend on tree topology except for exact-t ties.
#ifndef B2PT_BVH_H
#define B2PT_BVH_H

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <vector>

#include "b2pt_types.h"

namespace b2pt
{

struct BvhItem
{
  float bmin[3], bmax[3], cen[3];
  int32_t enc; // >=0 quad index, <0 ~sphere index
};

inline void quad_aabb(const B2Quad& Q, float* bmin, float* bmax)
{
  // vertices: q=v00, r=q+e01, s=v11, t=q+e03 would re-round; use the stored exact vertices where available
  float v[4][3];
  for (int c = 0; c < 3; ++c)
  {
    v[0][c] = Q.v00[c];
    v[2][c] = Q.v11[c];
    v[1][c] = Q.v11[c] + Q.e21[c]; // r = s + (r - s); exact when r-s was exact, otherwise within 1 ulp (padded below)
    v[3][c] = Q.v11[c] + Q.e23[c];
  }
  for (int c = 0; c < 3; ++c)
  {
    float lo = v[0][c], hi = v[0][c];
    for (int k = 1; k < 4; ++k)
    {
      lo = std::fmin(lo, v[k][c]);
      hi = std::fmax(hi, v[k][c]);
    }
    const float eps = std::fmax(1e-6f, 1.0e-4f * (hi - lo)); // AABBSurface.h:66-77
    bmin[c] = lo - eps;
    bmax[c] = hi + eps;
  }
}

inline void sphere_aabb(const B2Sphere& S, float* bmin, float* bmax)
{
  for (int c = 0; c < 3; ++c)
  {
    const float lo = S.c[c] - S.r, hi = S.c[c] + S.r;
    const float eps = std::fmax(1e-6f, 1.0e-4f * (hi - lo));
    bmin[c] = lo - eps;
    bmax[c] = hi + eps;
  }
}

inline float half_area(const float* bmin, const float* bmax)
{
  const float dx = bmax[0] - bmin[0], dy = bmax[1] - bmin[1], dz = bmax[2] - bmin[2];
  return dx * dy + dy * dz + dz * dx;
}

// Returns false if the tree does not fit the traversal's 24-bit index packing.
// The recursion is task-parallel (OpenMP, when compiled with -fopenmp): a node's two subtrees work on disjoint
// ranges of `items`; node pairs and leaf slot ranges are taken from atomic counters, so the numbering of nodes and
// slots depends on scheduling but the tree (boxes, splits, leaf contents and their order) does not.
struct BvhBuilder
{
  static constexpr int kBins = 16;
#ifndef B2PT_LEAF_TARGET
#define B2PT_LEAF_TARGET 4
#endif
#ifndef B2PT_LEAF_MAX
#define B2PT_LEAF_MAX 8
#endif
  static constexpr int kLeafTarget = B2PT_LEAF_TARGET;
  static constexpr int kLeafMax = B2PT_LEAF_MAX;
  static constexpr size_t kTaskMin = 8192; // subtrees below this size are built by the thread that reached them

  std::vector<BvhItem> items;
  std::vector<B2BvhNode>& nodes;
  std::vector<int32_t>& slots;
  std::atomic<int32_t> nodeCount{ 0 }, slotCount{ 0 }, maxDepth{ 0 };
  // SAH splits may peel one primitive per level on skewed input; past this depth every split is the balanced median
  // one, so the tree is never deeper than kSahDepthLimit + ceil(log2 n) <= 32 + 24 levels -- inside the traversal's
  // stack (64 entries for the binary layout)
  static constexpr int kSahDepthLimit = 32;

  BvhBuilder(std::vector<B2BvhNode>& n, std::vector<int32_t>& s)
    : nodes(n)
    , slots(s)
  {
  }
  static int64_t order_key(int32_t enc) { return enc >= 0 ? (int64_t)enc : ((int64_t)1 << 32) + (int64_t)(~enc); }

  void make_leaf(size_t lo, size_t hi, int32_t node)
  {
    std::sort(items.begin() + (std::ptrdiff_t)lo, items.begin() + (std::ptrdiff_t)hi,
              [](const BvhItem& a, const BvhItem& b) { return order_key(a.enc) < order_key(b.enc); });
    const int32_t n = (int32_t)(hi - lo);
    const int32_t first = slotCount.fetch_add(n);
    nodes[(size_t)node].left = first;
    nodes[(size_t)node].count = n;
    for (size_t i = lo; i < hi; ++i)
      slots[(size_t)first + (i - lo)] = items[i].enc;
  }

  void build(size_t wlo, size_t whi, int32_t wnode, int depth = 0)
  {
    const size_t n = whi - wlo;
    for (int32_t seen = maxDepth.load(); depth > seen && !maxDepth.compare_exchange_weak(seen, depth);)
    {
    }
    float bmin[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, bmax[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    float cmin[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, cmax[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (size_t i = wlo; i < whi; ++i)
      for (int c = 0; c < 3; ++c)
      {
        bmin[c] = std::fmin(bmin[c], items[i].bmin[c]);
        bmax[c] = std::fmax(bmax[c], items[i].bmax[c]);
        cmin[c] = std::fmin(cmin[c], items[i].cen[c]);
        cmax[c] = std::fmax(cmax[c], items[i].cen[c]);
      }
    for (int c = 0; c < 3; ++c)
    {
      nodes[(size_t)wnode].bmin[c] = bmin[c];
      nodes[(size_t)wnode].bmax[c] = bmax[c];
    }
    if (n <= 1)
      return make_leaf(wlo, whi, wnode);
    // binned SAH over the three axes
    int bestAxis = -1, bestSplit = -1;
    float bestCost = FLT_MAX;
    for (int axis = 0; axis < 3; ++axis)
    {
      const float ext = cmax[axis] - cmin[axis];
      if (!(ext > 0.f))
        continue;
      int cnt[kBins] = {};
      float bbmin[kBins][3], bbmax[kBins][3];
      for (int b = 0; b < kBins; ++b)
        for (int c = 0; c < 3; ++c)
        {
          bbmin[b][c] = FLT_MAX;
          bbmax[b][c] = -FLT_MAX;
        }
      const float scale = (float)kBins / ext;
      for (size_t i = wlo; i < whi; ++i)
      {
        int b = (int)((items[i].cen[axis] - cmin[axis]) * scale);
        b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
        cnt[b]++;
        for (int c = 0; c < 3; ++c)
        {
          bbmin[b][c] = std::fmin(bbmin[b][c], items[i].bmin[c]);
          bbmax[b][c] = std::fmax(bbmax[b][c], items[i].bmax[c]);
        }
      }
      float rightArea[kBins];
      int rightCnt[kBins];
      {
        float rmin[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, rmax[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
        int rc = 0;
        for (int b = kBins - 1; b >= 1; --b)
        {
          if (cnt[b])
            for (int c = 0; c < 3; ++c)
            {
              rmin[c] = std::fmin(rmin[c], bbmin[b][c]);
              rmax[c] = std::fmax(rmax[c], bbmax[b][c]);
            }
          rc += cnt[b];
          rightCnt[b] = rc;
          rightArea[b] = rc ? half_area(rmin, rmax) : 0.f;
        }
      }
      float lmin[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, lmax[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
      int lc = 0;
      for (int b = 0; b < kBins - 1; ++b)
      {
        if (cnt[b])
          for (int c = 0; c < 3; ++c)
          {
            lmin[c] = std::fmin(lmin[c], bbmin[b][c]);
            lmax[c] = std::fmax(lmax[c], bbmax[b][c]);
          }
        lc += cnt[b];
        if (lc == 0 || rightCnt[b + 1] == 0)
          continue;
        const float cost = half_area(lmin, lmax) * (float)lc + rightArea[b + 1] * (float)rightCnt[b + 1];
        if (cost < bestCost)
        {
          bestCost = cost;
          bestAxis = axis;
          bestSplit = b;
        }
      }
    }
    const float parentArea = half_area(bmin, bmax);
#ifndef B2PT_LEAF_COST
#define B2PT_LEAF_COST 1.0f
#endif
    const float leafCost = B2PT_LEAF_COST * (float)n; // intersection cost per primitive, in traversal steps
    const float splitCost = bestAxis >= 0 ? 1.5f + bestCost / parentArea : FLT_MAX; // traversal step ~1.5
    if (n <= (size_t)kLeafTarget && (splitCost >= leafCost || bestAxis < 0))
      return make_leaf(wlo, whi, wnode);
    if (bestAxis < 0 && n <= (size_t)kLeafMax)
      return make_leaf(wlo, whi, wnode);
    size_t mid;
    if (depth >= kSahDepthLimit && n > (size_t)kLeafMax)
      mid = wlo; // too deep: the median fallback below
    else if (bestAxis >= 0 && (splitCost < leafCost || n > (size_t)kLeafMax))
    {
      const float ext = cmax[bestAxis] - cmin[bestAxis];
      const float scale = (float)kBins / ext;
      const float lo = cmin[bestAxis];
      auto it = std::partition(items.begin() + (std::ptrdiff_t)wlo, items.begin() + (std::ptrdiff_t)whi,
                               [&](const BvhItem& a) {
                                 int b = (int)((a.cen[bestAxis] - lo) * scale);
                                 b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                                 return b <= bestSplit;
                               });
      mid = (size_t)(it - items.begin());
    }
    else if (n <= (size_t)kLeafMax)
      return make_leaf(wlo, whi, wnode);
    else
      mid = wlo; // force the median fallback below
    if (mid == wlo || mid == whi)
    { // coincident centroids: median split on primitive order
      mid = wlo + n / 2;
      std::nth_element(items.begin() + (std::ptrdiff_t)wlo, items.begin() + (std::ptrdiff_t)mid,
                       items.begin() + (std::ptrdiff_t)whi,
                       [](const BvhItem& a, const BvhItem& b) { return order_key(a.enc) < order_key(b.enc); });
    }
    const int32_t left = nodeCount.fetch_add(2); // sibling pair, adjacent in memory
    nodes[(size_t)wnode].left = left;
    nodes[(size_t)wnode].count = 0;
    if (n >= kTaskMin)
    {
#pragma omp task default(shared) firstprivate(wlo, mid, left, depth)
      build(wlo, mid, left, depth + 1);
      build(mid, whi, left + 1, depth + 1);
#pragma omp taskwait
    }
    else
    {
      build(wlo, mid, left, depth + 1);
      build(mid, whi, left + 1, depth + 1);
    }
  }
};

inline bool build_bvh(const std::vector<B2Quad>& quads, const std::vector<int32_t>& keptQuads,
                      const std::vector<B2Sphere>& sph, std::vector<B2BvhNode>& nodes, std::vector<int32_t>& slots,
                      int* depthOut = nullptr)
{
  BvhBuilder B(nodes, slots);
  std::vector<BvhItem>& items = B.items;
  items.reserve(keptQuads.size() + sph.size());
  for (int32_t q : keptQuads)
  {
    BvhItem it;
    quad_aabb(quads[(size_t)q], it.bmin, it.bmax);
    it.enc = q;
    items.push_back(it);
  }
  for (size_t s = 0; s < sph.size(); ++s)
  {
    BvhItem it;
    sphere_aabb(sph[s], it.bmin, it.bmax);
    it.enc = ~(int32_t)s;
    items.push_back(it);
  }
  for (auto& it : items)
    for (int c = 0; c < 3; ++c)
      it.cen[c] = 0.5f * (it.bmin[c] + it.bmax[c]);

  nodes.clear();
  slots.clear();
  if (items.empty())
    return true; // nothing for the tree (all primitives gated): traversal is skipped when nNodes == 0
  nodes.assign(2 * items.size() + 2, B2BvhNode{}); // a binary tree over n >= 1 leaves has at most 2n-1 nodes
  slots.assign(items.size(), 0);
  B.nodeCount = 2; // node 0 = root, node 1 unused: sibling pairs start at even indices = 64-byte aligned fetches
#pragma omp parallel
#pragma omp single nowait
  B.build(0, items.size(), 0);
  nodes.resize((size_t)B.nodeCount.load());
  if (depthOut)
    *depthOut = B.maxDepth.load() + 1;
  return nodes.size() < ((size_t)1 << 24) && slots.size() < ((size_t)1 << 24);
}

} // namespace b2pt
#endif
