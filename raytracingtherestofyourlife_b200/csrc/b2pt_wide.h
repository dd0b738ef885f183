// b2pt_wide.h -- host-side collapse of a binary BVH (b2pt_bvh.h SAH tree or the device LBVH) into the 8-wide
// compressed tree the traversal kernels walk (B2WideNode, b2pt_types.h).
//
// Replaces the flat binary LinearBVH the reference traverses (pathtracing/BVHTraverser.h:128-227, built by VTK-m through
// pathtracing/QuadIntersector.cxx:134 / SphereIntersector.cxx:75) for scenes too large for the kernel-parameter path.
// One 80-byte fetch brings eight child boxes (8-bit quantised against the node's own frame) instead of one 64-byte fetch
// per two: a third of the dependent memory round trips of the binary tree.  Layout and traversal order follow the
// compressed wide BVH of Ylitie, Karras & Laine (HPG 2017), restated for this code base:
//  * children are dealt to the eight slots so that visiting slots in ascending (slot XOR ray octant) order is
//    approximately front to back -- no per-node sorting in the kernel;
//  * inner children are stored contiguously (child index = childBase + rank among the inner slots), primitive children
//    are single primitives whose slots in the leaf-ordered arrays are contiguous too (primBase + rank): a "leaf" is a
//    node whose children are primitives, and the same box test that culls subtrees culls individual primitives
//    before their exact (reference-arithmetic) test runs.
// Quantisation is conservative with a margin: boxes are rounded outwards and then widened by one more step on every
// side, which covers the half-step error of the kernel's magic-number dequantisation (b2pt_device.cuh wide_node_hits)
// and the few-ulp differences between the culling arithmetic (FMA) and the exact primitive tests.  Culling volumes never
// decide a hit: ids and t come from the exact tests, so results are those of brute force (tests/test_gpu_spheres.py).
#ifndef B2PT_WIDE_H
#define B2PT_WIDE_H

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "b2pt_bvh.h"
#include "b2pt_types.h"

namespace b2pt
{

struct WideBuildResult
{
  std::vector<B2WideNode> nodes; // node 0 = root
  std::vector<int32_t> slots;    // leaf order -> encoded primitive (>= 0 quad index, < 0 ~sphere index)
  int maxDepth = 0;              // levels of wide nodes (root = 1)
};

namespace wide_detail
{
struct Cand
{
  int32_t bin; // binary node index, or -1 for a single primitive
  int32_t enc; // primitive (bin == -1)
  float lo[3], hi[3];
  float area() const { return half_area(lo, hi); }
};

inline void prim_box(const std::vector<B2Quad>& quads, const std::vector<B2Sphere>& sph, int32_t enc, float* lo, float* hi)
{
  if (enc >= 0)
    quad_aabb(quads[(size_t)enc], lo, hi);
  else
    sphere_aabb(sph[(size_t)(~enc)], lo, hi);
}
} // namespace wide_detail

// binNodes / binSlots: the binary tree (B2BvhNode: count == 0 inner with children left, left+1; count > 0 leaf of
// binSlots[left .. left+count)).  sceneAbs: max |coordinate| (scale of the extra padding).
inline bool collapse_to_wide(const std::vector<B2BvhNode>& binNodes, const std::vector<int32_t>& binSlots,
                             const std::vector<B2Quad>& quads, const std::vector<B2Sphere>& sph, float sceneAbs,
                             WideBuildResult& out)
{
  using wide_detail::Cand;
  out.nodes.clear();
  out.slots.clear();
  out.maxDepth = 0;
  if (binNodes.empty())
    return true;
  struct Work
  {
    int32_t bin;
    int32_t wide;
    int depth;
  };
  std::vector<Work> queue;
  out.nodes.emplace_back();
  queue.push_back({ 0, 0, 1 });
  const double extraPad = 1e-6 * std::max(1.0, (double)sceneAbs);
  auto expand = [&](const Cand& c, std::vector<Cand>& dst) {
    const B2BvhNode& n = binNodes[(size_t)c.bin];
    if (n.count > 0)
      for (int k = 0; k < n.count; ++k)
      {
        Cand p;
        p.bin = -1;
        p.enc = binSlots[(size_t)n.left + k];
        wide_detail::prim_box(quads, sph, p.enc, p.lo, p.hi);
        dst.push_back(p);
      }
    else
      for (int k = 0; k < 2; ++k)
      {
        const B2BvhNode& ch = binNodes[(size_t)n.left + k];
        Cand p;
        p.bin = n.left + k;
        p.enc = 0;
        for (int a = 0; a < 3; ++a)
          p.lo[a] = ch.bmin[a], p.hi[a] = ch.bmax[a];
        dst.push_back(p);
      }
  };
  auto expansion_size = [&](const Cand& c) {
    if (c.bin < 0)
      return 1;
    const int n = binNodes[(size_t)c.bin].count;
    return n > 0 ? n : 2;
  };
  for (size_t qi = 0; qi < queue.size(); ++qi)
  {
    const Work w = queue[qi];
    out.maxDepth = std::max(out.maxDepth, w.depth);
    std::vector<Cand> cands;
    {
      Cand root;
      root.bin = w.bin;
      root.enc = 0;
      const B2BvhNode& n = binNodes[(size_t)w.bin];
      if (n.count > 8)
        return false; // (the builders emit leaves of at most 8 primitives)
      for (int a = 0; a < 3; ++a)
        root.lo[a] = n.bmin[a], root.hi[a] = n.bmax[a];
      expand(root, cands);
    }
    for (;;)
    { // greedy: open the largest candidate that still fits
      int best = -1;
      float bestArea = -1.f;
      for (size_t i = 0; i < cands.size(); ++i)
        if (cands[i].bin >= 0 && (int)cands.size() - 1 + expansion_size(cands[i]) <= 8 && cands[i].area() > bestArea)
          best = (int)i, bestArea = cands[i].area();
      if (best < 0)
        break;
      const Cand c = cands[(size_t)best];
      cands.erase(cands.begin() + best);
      expand(c, cands);
    }
    // node frame
    double lo[3] = { DBL_MAX, DBL_MAX, DBL_MAX }, hi[3] = { -DBL_MAX, -DBL_MAX, -DBL_MAX };
    for (const Cand& c : cands)
      for (int a = 0; a < 3; ++a)
      {
        lo[a] = std::min(lo[a], (double)c.lo[a] - extraPad);
        hi[a] = std::max(hi[a], (double)c.hi[a] + extraPad);
      }
    B2WideNode N;
    std::memset(&N, 0, sizeof(N));
    double scale[3];
    uint8_t* ex[3] = { &N.ex, &N.ey, &N.ez };
    for (int a = 0; a < 3; ++a)
    {
      // steps of 2^e: 252 of them span the node (two are kept free at either end for the outward widening), and a step
      // is never finer than 2^-18 of the coordinates' magnitude (float32 resolution of the frame origin is 2^-24)
      const double ext = std::max(hi[a] - lo[a], 0.0);
      const double mag = std::max(std::max(std::fabs(lo[a]), std::fabs(hi[a])), 1e-30);
      int e = (int)std::ceil(std::log2(std::max(ext / 252.0, mag * std::ldexp(1.0, -18))));
      e = std::max(-100, std::min(100, e));
      for (;;)
      {
        scale[a] = std::ldexp(1.0, e);
        N.org[a] = (float)(lo[a] - 2.0 * scale[a]);
        // re-check against the rounded origin (the float origin may sit a little above the double one)
        if ((hi[a] - (double)N.org[a]) / scale[a] <= 254.0 && (lo[a] - (double)N.org[a]) / scale[a] >= 1.0)
          break;
        ++e;
      }
      *ex[a] = (uint8_t)(e + 127);
    }
    // slots: greedy assignment of children to octant corners (score = centroid offset along the slot's diagonal)
    double cen[3] = { 0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2]) };
    int slotOf[8];
    bool slotUsed[8] = {}, candDone[8] = {};
    for (size_t it = 0; it < cands.size(); ++it)
    {
      double bestScore = -DBL_MAX;
      int bc = -1, bs = -1;
      for (size_t c = 0; c < cands.size(); ++c)
      {
        if (candDone[c])
          continue;
        for (int s = 0; s < 8; ++s)
        {
          if (slotUsed[s])
            continue;
          double score = 0.0;
          for (int a = 0; a < 3; ++a)
          {
            const double off = 0.5 * ((double)cands[c].lo[a] + (double)cands[c].hi[a]) - cen[a];
            score += ((s >> a) & 1) ? off : -off;
          }
          if (score > bestScore)
            bestScore = score, bc = (int)c, bs = s;
        }
      }
      candDone[bc] = true;
      slotUsed[bs] = true;
      slotOf[bc] = bs;
    }
    for (int s = 0; s < 8; ++s)
      for (int a = 0; a < 3; ++a)
        N.qlo[a][s] = 255, N.qhi[a][s] = 0; // empty slot: inverted box
    int candAt[8];
    std::fill(candAt, candAt + 8, -1);
    for (size_t c = 0; c < cands.size(); ++c)
      candAt[slotOf[c]] = (int)c;
    N.childBase = (uint32_t)out.nodes.size();
    N.primBase = (uint32_t)out.slots.size();
    for (int s = 0; s < 8; ++s)
    {
      if (candAt[s] < 0)
        continue;
      const Cand& c = cands[(size_t)candAt[s]];
      for (int a = 0; a < 3; ++a)
      {
        const double ql = std::floor(((double)c.lo[a] - extraPad - (double)N.org[a]) / scale[a]) - 1.0;
        const double qh = std::ceil(((double)c.hi[a] + extraPad - (double)N.org[a]) / scale[a]) + 1.0;
        if (!(ql >= 0.0 && qh <= 255.0))
          return false; // cannot happen (frame construction above)
        N.qlo[a][s] = (uint8_t)ql;
        N.qhi[a][s] = (uint8_t)qh;
      }
      if (c.bin < 0)
      {
        N.primMask |= (uint8_t)(1u << s);
        out.slots.push_back(c.enc);
      }
      else
      {
        N.innerMask |= (uint8_t)(1u << s);
        out.nodes.emplace_back();
        queue.push_back({ c.bin, (int32_t)out.nodes.size() - 1, w.depth + 1 });
      }
    }
    out.nodes[(size_t)w.wide] = N;
  }
  return out.nodes.size() < ((size_t)1 << 31) && out.slots.size() < ((size_t)1 << 31);
}

// Debug self-check (B2PT_VALIDATE_BVH): every primitive is reachable exactly once, and its padded box lies inside the
// dequantised box of its slot in every node on the way down, with one quantisation step to spare on each side.
inline bool validate_wide(const WideBuildResult& W, const std::vector<B2Quad>& quads, const std::vector<B2Sphere>& sph,
                          std::string& why)
{
  std::vector<int> seen(W.slots.size(), 0);
  struct Item
  {
    uint32_t node;
    double lo[3], hi[3];
  };
  std::vector<Item> st;
  Item r;
  r.node = 0;
  for (int a = 0; a < 3; ++a)
    r.lo[a] = -DBL_MAX, r.hi[a] = DBL_MAX;
  st.push_back(r);
  size_t visited = 0;
  while (!st.empty())
  {
    const Item it = st.back();
    st.pop_back();
    if (it.node >= W.nodes.size())
    {
      why = "child index out of range";
      return false;
    }
    const B2WideNode& N = W.nodes[it.node];
    if (N.innerMask & N.primMask)
    {
      why = "slot is both inner and primitive";
      return false;
    }
    const uint8_t ex[3] = { N.ex, N.ey, N.ez };
    uint32_t innerRank = 0, primRank = 0;
    for (int s = 0; s < 8; ++s)
    {
      const bool inner = (N.innerMask >> s) & 1, prim = (N.primMask >> s) & 1;
      if (!inner && !prim)
        continue;
      Item ch;
      for (int a = 0; a < 3; ++a)
      {
        const double sc = std::ldexp(1.0, (int)ex[a] - 127);
        // the box the kernel culls with, shrunk by the step the dequantisation may be off by
        ch.lo[a] = (double)N.org[a] + sc * ((double)N.qlo[a][s] + 1.0);
        ch.hi[a] = (double)N.org[a] + sc * ((double)N.qhi[a][s] - 1.0);
        if (ch.lo[a] < it.lo[a] - 3.0 * sc || ch.hi[a] > it.hi[a] + 3.0 * sc)
        { // (a child frame may stick out of its parent's quantised box by rounding, never by more than a few steps)
          why = "child box outside its parent";
          return false;
        }
      }
      if (inner)
      {
        ch.node = N.childBase + innerRank++;
        st.push_back(ch);
      }
      else
      {
        const size_t slot = (size_t)N.primBase + primRank++;
        if (slot >= W.slots.size())
        {
          why = "primitive slot out of range";
          return false;
        }
        ++seen[slot];
        ++visited;
        float lo[3], hi[3];
        wide_detail::prim_box(quads, sph, W.slots[slot], lo, hi);
        for (int a = 0; a < 3; ++a)
          if ((double)lo[a] < ch.lo[a] || (double)hi[a] > ch.hi[a])
          {
            why = "primitive box outside its quantised box";
            return false;
          }
      }
    }
  }
  for (int v : seen)
    if (v != 1)
    {
      why = "a primitive slot is not reached exactly once";
      return false;
    }
  return visited == W.slots.size();
}

} // namespace b2pt
#endif
