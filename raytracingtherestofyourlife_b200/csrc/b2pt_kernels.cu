// b2pt_kernels.cu -- hand-written sm_100a kernels of the wavefront path tracer.
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo (see _build.py).
//
//   k_bounce<PRIMARY,Scene>  one launch per bounce.  PRIMARY=true fuses per-pixel camera-ray generation
//                            (wang-seeded RNG) into bounce 0, so primary rays never touch HBM.  Every
//                            launch: load ray (3x16 B, coalesced) -> closest hit -> fused material /
//                            direction sampling / mixture pdf / emission -> survivors compacted with
//                            warp ballot + block prefix + ONE atomicAdd per 256-ray tile into the next
//                            SoA queue.  Terminated paths write their radiance once (16 B).
//   k_accumulate             per pixel: sum the batch's samples in registers (sample order) and do a
//                            single read-modify-write of the canvas float4.
//   k_primary_hits           parity hook: production raygen + trace, writes hit primitive id / t.
//   k_create_rays, k_intersect, k_normalize, k_fill_seeds, k_sum_peers: stage-level API kernels.
//
// Grid sizing: persistent CTAs, numSMs x occupancy blocks, each striding over 256-ray tiles; the input
// count of bounce d>0 lives in device memory (counters[d-1]) so no host round trip sits between bounces.
#include "b2pt_device.cuh"
#include "b2pt_kernels.h"

namespace b2pt
{

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;

struct PeerPtrs
{
  const float4* p[8];
};

__device__ __forceinline__ void store_ray(const B2Queue& q, uint32_t pos, f3 o, f3 d, f3 T, uint32_t pid, uint32_t rng)
{
  q.p0[pos] = make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(d.x));
  q.p1[pos] = make_uint4(__float_as_uint(d.y), __float_as_uint(d.z), __float_as_uint(T.x), __float_as_uint(T.y));
  q.p2[pos] = make_uint4(__float_as_uint(T.z), pid, rng, 0u);
}

template <bool PRIMARY, class SceneT>
__global__ void __launch_bounds__(kBlock)
  k_bounce(const __grid_constant__ B2Camera cam, const __grid_constant__ SceneT scene,
           const __grid_constant__ B2Lights lights, const __grid_constant__ B2RenderArgs A)
{
  __shared__ uint32_t sWarp[kWarps];
  __shared__ uint32_t sBase;

  const int64_t nIn = PRIMARY ? A.nPaths : (int64_t)A.counters[A.depth - 1];
  const bool lastDepth = (A.depth == A.maxDepth - 1);
  const bool refStream = (A.flags & B2PT_FLAG_REFERENCE_STREAM_DEV) != 0;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;

  for (int64_t tile = blockIdx.x; tile * kBlock < nIn; tile += gridDim.x)
  {
    const int64_t idx = tile * kBlock + threadIdx.x;
    bool survive = false;
    f3 o, d, T;
    uint32_t pid = 0, rng = 0;
    if (idx < nIn)
    {
      if (PRIMARY)
      {
        pid = (uint32_t)idx;
        const uint32_t pixel = pid % (uint32_t)A.nPixels;
        const uint32_t b = pid / (uint32_t)A.nPixels;
        // MapperPathTracer.cxx:265-267: seeds[i] = i.  Production stream: one stream per (pixel, sample).
        rng = refStream ? A.seeds[pixel] : pixel + A.seedOffset + (uint32_t)(A.sampleBase + (int)b) * B2PT_GOLDEN;
        d = raygen(cam, (int)pixel, rng);
        o = ld3(cam.pos);
        T = mk3(1.f, 1.f, 1.f);
      }
      else
      {
        const uint4 a = A.qin.p0[idx];
        const uint4 b = A.qin.p1[idx];
        const uint4 c = A.qin.p2[idx];
        o = mk3(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z));
        d = mk3(__uint_as_float(a.w), __uint_as_float(b.x), __uint_as_float(b.y));
        T = mk3(__uint_as_float(b.z), __uint_as_float(b.w), __uint_as_float(c.x));
        pid = c.y;
        rng = c.z;
      }
      f3 L;
      Hit hit;
      BounceResult r = bounce(scene, lights, o, d, T, rng, A.flags, L, hit);
      if (r == BOUNCE_CONTINUE && lastDepth)
      {
        L = T * 0.f; // still alive after maxDepth bounces: e[D-1] = 0 (MapperPathTracer.cxx:328-331)
        r = BOUNCE_DONE;
        if (refStream)
          A.seeds[pid % (uint32_t)A.nPixels] = rng;
      }
      else if (r == BOUNCE_DONE && refStream)
      {
        burn_depths(rng, A.maxDepth - A.depth);
        A.seeds[pid % (uint32_t)A.nPixels] = rng;
      }
      if (r == BOUNCE_DONE)
        A.rad[pid] = make_float4(L.x, L.y, L.z, 0.f);
      else
        survive = true;
    }
    // ---- K5: compaction of survivors into the next queue -------------------------------------
    const unsigned ballot = __ballot_sync(0xffffffffu, survive);
    if (lane == 0)
      sWarp[warp] = __popc(ballot);
    __syncthreads();
    if (threadIdx.x == 0)
    {
      uint32_t run = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w)
      {
        uint32_t c = sWarp[w];
        sWarp[w] = run;
        run += c;
      }
      sBase = run ? atomicAdd(&A.counters[A.depth], run) : 0u;
    }
    __syncthreads();
    if (survive)
    {
      const uint32_t pos = sBase + sWarp[warp] + __popc(ballot & ((1u << lane) - 1u));
      store_ray(A.qout, pos, o, d, T, pid, rng);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
  k_accumulate(float4* __restrict__ color, const float4* __restrict__ rad, int nPixels, int samplesInBatch,
               unsigned long long* nanCounter)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nPixels)
    return;
  float4 c = color[p];
  int nan = 0;
  for (int b = 0; b < samplesInBatch; ++b)
  {
    const float4 r = rad[(size_t)b * nPixels + p];
    nan += (r.x != r.x || r.y != r.y || r.z != r.z) ? 1 : 0;
    c.x += r.x; // cols += sumtotl, MapperPathTracer.cxx:350, in sample order
    c.y += r.y;
    c.z += r.z;
  }
  color[p] = c;
  if (nan)
    atomicAdd(nanCounter, (unsigned long long)nan);
}

template <class SceneT>
__global__ void __launch_bounds__(256)
  k_primary_hits(const __grid_constant__ B2Camera cam, const __grid_constant__ SceneT scene, uint32_t seedOffset,
                 int32_t* primOut, float* tOut)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= cam.W * cam.H)
    return;
  uint32_t rng = (uint32_t)p + seedOffset;
  f3 d = raygen(cam, p, rng);
  Hit hit;
  bool h = trace(scene, ld3(cam.pos), d, 0.001f, FLT_MAX, hit);
  if (primOut)
    primOut[p] = h ? hit.prim : -1;
  if (tOut)
    tOut[p] = h ? hit.t : FLT_MAX;
}

// pathtracing/Camera.cxx:894-953: fills + RayGen + origin broadcast, fused.
__global__ void __launch_bounds__(256)
  k_create_rays(const __grid_constant__ B2Camera cam, uint32_t* seeds, float* dx, float* dy, float* dz, float* ox,
                float* oy, float* oz, long long* pixelIdx)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= cam.W * cam.H)
    return;
  uint32_t rng = seeds[p];
  f3 d = raygen(cam, p, rng);
  seeds[p] = rng;
  if (dx)
    dx[p] = d.x, dy[p] = d.y, dz[p] = d.z;
  if (ox)
    ox[p] = cam.pos[0], oy[p] = cam.pos[1], oz[p] = cam.pos[2];
  if (pixelIdx)
    pixelIdx[p] = p;
}

template <class SceneT>
__global__ void __launch_bounds__(256)
  k_intersect(const __grid_constant__ SceneT scene, long long n, const float* ox, const float* oy, const float* oz,
              const float* dx, const float* dy, const float* dz, float tmin, float tmax, int32_t* primId, float* hrec9,
              int32_t* matId, int32_t* texId)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  Hit hit;
  bool h = trace(scene, mk3(ox[i], oy[i], oz[i]), mk3(dx[i], dy[i], dz[i]), tmin, tmax, hit);
  primId[i] = h ? hit.prim : -1;
  if (matId)
    matId[i] = h ? hit.mat : -1;
  if (texId)
    texId[i] = h ? hit.texi : -1;
  if (hrec9)
  {
    // Record.h:4 order U,V,T,Nx,Ny,Nz,Px,Py,Pz; u,v are never consumed by the path tracer: written as 0
    hrec9[0 * n + i] = 0.f;
    hrec9[1 * n + i] = 0.f;
    hrec9[2 * n + i] = h ? hit.t : tmax;
    hrec9[3 * n + i] = h ? hit.n.x : 0.f;
    hrec9[4 * n + i] = h ? hit.n.y : 0.f;
    hrec9[5 * n + i] = h ? hit.n.z : 0.f;
    hrec9[6 * n + i] = h ? hit.p.x : 0.f;
    hrec9[7 * n + i] = h ? hit.p.y : 0.f;
    hrec9[8 * n + i] = h ? hit.p.z : 0.f;
  }
}

// main.cc:253-287 NormalizeFunctor: sqrt(de_nan(sum)/spp); the alpha lane goes through the same sqrt.
__global__ void __launch_bounds__(256) k_normalize(float4* color, long long n, float spp)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  float4 c = color[i];
  if (!(c.x == c.x))
    c.x = 0.f;
  if (!(c.y == c.y))
    c.y = 0.f;
  if (!(c.z == c.z))
    c.z = 0.f;
  color[i] = make_float4(sqrtf(c.x / spp), sqrtf(c.y / spp), sqrtf(c.z / spp), sqrtf(c.w / spp));
}

__global__ void __launch_bounds__(256) k_fill_seeds(uint32_t* seeds, int n, uint32_t seedOffset)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    seeds[i] = (uint32_t)i + seedOffset; // MapperPathTracer.cxx:265-267
}

__global__ void __launch_bounds__(256)
  k_sum_peers(float4* dst, const __grid_constant__ PeerPtrs srcs, int G, long long begin, long long end)
{
  for (long long i = begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < end;
       i += (long long)gridDim.x * blockDim.x)
  {
    float4 acc = srcs.p[0][i];
    for (int g = 1; g < G; ++g)
    {
      const float4 v = srcs.p[g][i];
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    dst[i - begin] = acc;
  }
}

// --------------------------------------------------------------------------------------- launchers
cudaError_t query_launch_cfg(LaunchCfg* cfg)
{
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess)
    return e;
  e = cudaDeviceGetAttribute(&cfg->numSMs, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess)
    return e;
  int n = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_bounce<true, B2SmallScene>, kBlock, 0);
  if (e != cudaSuccess)
    return e;
  cfg->bounceBlocksPerSM[1][0] = n > 0 ? n : 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_bounce<false, B2SmallScene>, kBlock, 0);
  if (e != cudaSuccess)
    return e;
  cfg->bounceBlocksPerSM[0][0] = n > 0 ? n : 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_bounce<true, B2BvhScene>, kBlock, 0);
  if (e != cudaSuccess)
    return e;
  cfg->bounceBlocksPerSM[1][1] = n > 0 ? n : 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_bounce<false, B2BvhScene>, kBlock, 0);
  if (e != cudaSuccess)
    return e;
  cfg->bounceBlocksPerSM[0][1] = n > 0 ? n : 1;
  return cudaSuccess;
}

cudaError_t launch_bounce(const LaunchCfg& cfg, bool primary, const B2Camera& cam, const B2SmallScene* small,
                          const B2BvhScene* bvh, const B2Lights& lights, const B2RenderArgs& args, int64_t maxRaysIn,
                          cudaStream_t stream)
{
  const int useBvh = bvh ? 1 : 0;
  int64_t tiles = (maxRaysIn + kBlock - 1) / kBlock;
  int64_t persistent = (int64_t)cfg.numSMs * cfg.bounceBlocksPerSM[primary ? 1 : 0][useBvh];
  int grid = (int)(tiles < persistent ? tiles : persistent);
  if (grid < 1)
    grid = 1;
  if (bvh)
  {
    if (primary)
      k_bounce<true, B2BvhScene><<<grid, kBlock, 0, stream>>>(cam, *bvh, lights, args);
    else
      k_bounce<false, B2BvhScene><<<grid, kBlock, 0, stream>>>(cam, *bvh, lights, args);
  }
  else
  {
    if (primary)
      k_bounce<true, B2SmallScene><<<grid, kBlock, 0, stream>>>(cam, *small, lights, args);
    else
      k_bounce<false, B2SmallScene><<<grid, kBlock, 0, stream>>>(cam, *small, lights, args);
  }
  return cudaGetLastError();
}

cudaError_t launch_accumulate(float4* color, const float4* rad, int nPixels, int samplesInBatch,
                              unsigned long long* nanCounter, cudaStream_t stream)
{
  k_accumulate<<<(nPixels + 255) / 256, 256, 0, stream>>>(color, rad, nPixels, samplesInBatch, nanCounter);
  return cudaGetLastError();
}

cudaError_t launch_primary_hits(const B2Camera& cam, const B2SmallScene* small, const B2BvhScene* bvh,
                                uint32_t seedOffset, int32_t* primOut, float* tOut, cudaStream_t stream)
{
  const int n = cam.W * cam.H;
  if (bvh)
    k_primary_hits<B2BvhScene><<<(n + 255) / 256, 256, 0, stream>>>(cam, *bvh, seedOffset, primOut, tOut);
  else
    k_primary_hits<B2SmallScene><<<(n + 255) / 256, 256, 0, stream>>>(cam, *small, seedOffset, primOut, tOut);
  return cudaGetLastError();
}

cudaError_t launch_create_rays(const B2Camera& cam, uint32_t* seeds, float* dx, float* dy, float* dz, float* ox,
                               float* oy, float* oz, long long* pixelIdx, cudaStream_t stream)
{
  const int n = cam.W * cam.H;
  k_create_rays<<<(n + 255) / 256, 256, 0, stream>>>(cam, seeds, dx, dy, dz, ox, oy, oz, pixelIdx);
  return cudaGetLastError();
}

cudaError_t launch_intersect(const B2SmallScene* small, const B2BvhScene* bvh, int64_t n, const float* ox,
                             const float* oy, const float* oz, const float* dx, const float* dy, const float* dz,
                             float tmin, float tmax, int32_t* primId, float* hrec9, int32_t* matId, int32_t* texId,
                             cudaStream_t stream)
{
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (grid == 0)
    return cudaSuccess;
  if (bvh)
    k_intersect<B2BvhScene>
      <<<grid, 256, 0, stream>>>(*bvh, n, ox, oy, oz, dx, dy, dz, tmin, tmax, primId, hrec9, matId, texId);
  else
    k_intersect<B2SmallScene>
      <<<grid, 256, 0, stream>>>(*small, n, ox, oy, oz, dx, dy, dz, tmin, tmax, primId, hrec9, matId, texId);
  return cudaGetLastError();
}

cudaError_t launch_normalize(float4* color, int64_t n, int spp, cudaStream_t stream)
{
  if (n <= 0)
    return cudaSuccess;
  k_normalize<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(color, n, (float)spp);
  return cudaGetLastError();
}

cudaError_t launch_fill_seeds(uint32_t* seeds, int n, uint32_t seedOffset, cudaStream_t stream)
{
  if (n <= 0)
    return cudaSuccess;
  k_fill_seeds<<<(n + 255) / 256, 256, 0, stream>>>(seeds, n, seedOffset);
  return cudaGetLastError();
}

cudaError_t launch_sum_peers(float4* dst, const float4* const* srcs, int G, int64_t begin, int64_t end,
                             cudaStream_t stream)
{
  if (G < 1 || G > 8)
    return cudaErrorInvalidValue;
  if (end <= begin)
    return cudaSuccess;
  PeerPtrs pp;
  for (int g = 0; g < 8; ++g)
    pp.p[g] = g < G ? srcs[g] : nullptr;
  int64_t n = end - begin;
  int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  k_sum_peers<<<grid, 256, 0, stream>>>(dst, pp, G, begin, end);
  return cudaGetLastError();
}

} // namespace b2pt
