// b2pt_lbvh.h -- GPU LBVH builder (implemented in b2pt_lbvh.cu).
#ifndef B2PT_LBVH_H
#define B2PT_LBVH_H

#include <cuda_runtime.h>

#include "b2pt_types.h"

namespace b2pt
{

// Builds the traversal structures of a BVH scene on the device: n = nTreeQuads + nSph primitives (quads
// dQuads[dTreeQuads[i]], then spheres dSph[0..nSph)).  Outputs, all preallocated by the caller:
//   dNodesOut   2*n B2BvhNode records (root at 0, children of radix node i at 2i+2 / 2i+3)
//   dSlotsOut   n encoded primitives in Morton order (>= 0 quad index, < 0 ~sphere index)
//   dLeafSphOut n leaf-ordered sphere geometries
// sceneLo/Hi: bounds of the primitive centroids' domain (Morton quantisation).  Synchronises `stream`.
cudaError_t build_lbvh_device(const B2Quad* dQuads, const B2Sphere* dSph, const int32_t* dTreeQuads, int nTreeQuads,
                              int nSph, const float sceneLo[3], const float sceneHi[3], B2BvhNode* dNodesOut,
                              int32_t* dSlotsOut, float4* dLeafSphOut, cudaStream_t stream);

// Bottom-up refit of an existing tree (either builder's B2BvhNode layout) to the primitives' current geometry.
// dParent[i]: parent node of node i, -1 for the root, -2 for entries that are not part of the tree; dArrived: nNodes
// scratch counters.  Also rewrites the leaf-ordered sphere geometry.  Asynchronous on `stream`.
cudaError_t refit_bvh_device(B2BvhNode* dNodes, const int32_t* dParent, int* dArrived, int nNodes,
                             const int32_t* dSlots, const B2Quad* dQuads, const B2Sphere* dSph, float4* dLeafSph,
                             cudaStream_t stream);

} // namespace b2pt
#endif
