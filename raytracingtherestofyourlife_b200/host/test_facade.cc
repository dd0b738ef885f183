// test_facade.cc -- C++ tests of the drop-in facade, written against the reference's class surface.
//   test_facade --cpu            host-only behaviour (containers, validation, error messages); no GPU needed
//   test_facade --multigpu G     MapperPathTracer::SetDevices on G GPUs of this box against the one-GPU render
//   test_facade --gpu <prefix>   renders through MapperPathTracer / Camera::CreateRays / intersect on cuda:0 and
//                                writes <prefix>_*.bin for tests/test_facade.py to compare with the oracle
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <string>
#include <vector>

#include "CornellBox.h"
#include "MapperPathTracer.h"
#include "b2pt_facade.h"
#include "pathtracing/Camera.h"
#include "pathtracing/PathTracer.h"
#include "raytracing/ChannelBufferOperations.h"

static int g_fail = 0;
#define CHECK(cond)                                                                                                    \
  do                                                                                                                   \
  {                                                                                                                    \
    if (!(cond))                                                                                                       \
    {                                                                                                                  \
      std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);                                                    \
      ++g_fail;                                                                                                        \
    }                                                                                                                  \
  } while (0)

static bool throwsBadValue(const std::function<void()>& f, const char* contains = nullptr)
{
  try
  {
    f();
  }
  catch (const vtkm::cont::ErrorBadValue& e)
  {
    return !contains || e.GetMessage().find(contains) != std::string::npos;
  }
  catch (...)
  {
    return false;
  }
  return false;
}

using CB = vtkm::rendering::raytracing::ChannelBuffer<vtkm::Float32>;
using RayF = vtkm::rendering::raytracing::Ray<vtkm::Float32>;

static void testChannelBuffer()
{
  CB a(3, 4);
  CHECK(a.GetNumChannels() == 3 && a.GetSize() == 4 && a.GetBufferLength() == 12 && a.GetName() == "default");
  a.InitConst(2.f);
  CB b = a.Copy();
  b.InitConst(0.5f);
  a.AddBuffer(b);
  CHECK(a.Buffer.ReadPortal().Get(7) == 2.5f);
  a.MultiplyBuffer(b);
  CHECK(a.Buffer.ReadPortal().Get(0) == 1.25f);
  vtkm::cont::ArrayHandle<vtkm::Float32> sig = vtkm::cont::make_ArrayHandle(std::vector<float>{ 1.f, 2.f, 3.f });
  a.InitChannels(sig);
  CHECK(a.Buffer.ReadPortal().Get(4) == 2.f && a.Buffer.ReadPortal().Get(11) == 3.f);
  CB ch = a.GetChannel(2);
  CHECK(ch.GetNumChannels() == 1 && ch.GetSize() == 4 && ch.Buffer.ReadPortal().Get(3) == 3.f);
  CHECK(throwsBadValue([&] { a.GetChannel(3); }, "invalid channel"));
  CHECK(throwsBadValue([&] { a.GetChannel(-1); }));
  CB c(2, 4);
  CHECK(throwsBadValue([&] { a.AddBuffer(c); }, "number of channels must be equal"));
  CB d(3, 5);
  CHECK(throwsBadValue([&] { a.MultiplyBuffer(d); }, "size must be equal"));
  CHECK(throwsBadValue([&] { a.Resize(-1); }, "Size must be greater than -1"));
  CHECK(throwsBadValue([&] { a.SetNumChannels(0); }));
  CHECK(throwsBadValue([] { CB bad(-1, 3); }));
  a.SetNumChannels(2);
  CHECK(a.GetBufferLength() == 8);
  a.Resize(0);
  CHECK(a.GetBufferLength() == 0); // empty buffers are legal
  CB e0(1, 0);
  CHECK(e0.GetChannel(0).GetSize() == 0);
  // normalise: (v-min)/(max-min), inverted
  CB n(1, 3);
  n.Buffer.WritePortal().Set(0, 2.f), n.Buffer.WritePortal().Set(1, 4.f), n.Buffer.WritePortal().Set(2, 6.f);
  n.Normalize(false);
  CHECK(n.Buffer.ReadPortal().Get(0) == 0.f && n.Buffer.ReadPortal().Get(1) == 0.5f && n.Buffer.ReadPortal().Get(2) == 1.f);
  n.Normalize(true);
  CHECK(n.Buffer.ReadPortal().Get(0) == 1.f && n.Buffer.ReadPortal().Get(2) == 0.f);
  // compact (scan + scatter) and expand (its inverse)
  CB q(2, 5);
  for (int i = 0; i < 10; ++i)
    q.Buffer.WritePortal().Set(i, float(i));
  vtkm::cont::ArrayHandle<vtkm::UInt8> mask = vtkm::cont::make_ArrayHandle(std::vector<vtkm::UInt8>{ 1, 0, 0, 1, 1 });
  vtkm::rendering::raytracing::ChannelBufferOperations::Compact(q, mask, 3);
  CHECK(q.GetSize() == 3 && q.Buffer.ReadPortal().Get(2) == 6.f && q.Buffer.ReadPortal().Get(5) == 9.f);
  vtkm::cont::ArrayHandle<vtkm::Id> sparse = vtkm::cont::make_ArrayHandle(std::vector<vtkm::Id>{ 0, 3, 4 });
  CB x = q.ExpandBuffer(sparse, 5, -1.f);
  CHECK(x.GetSize() == 5 && x.Buffer.ReadPortal().Get(2) == -1.f && x.Buffer.ReadPortal().Get(6) == 6.f &&
        x.Buffer.ReadPortal().Get(9) == 9.f);
  vtkm::cont::ArrayHandle<vtkm::Float32> sig2 = vtkm::cont::make_ArrayHandle(std::vector<float>{ 7.f, 8.f });
  CB y = q.ExpandBuffer(sparse, 5, sig2);
  CHECK(y.Buffer.ReadPortal().Get(2) == 7.f && y.Buffer.ReadPortal().Get(3) == 8.f && y.Buffer.ReadPortal().Get(8) == 8.f);
  CHECK(throwsBadValue([&] { vtkm::rendering::raytracing::ChannelBufferOperations::Compact(q, mask, 3); }));
}

static void testRay()
{
  RayF rays;
  CHECK(rays.NumRays == 0 && rays.Buffers.size() == 1);
  rays.AddBuffer(1, "sum_values");
  rays.AddBuffer(5, "attenuationX");
  rays.Resize(16);
  CHECK(rays.DirX.GetNumberOfValues() == 16 && rays.Status.GetNumberOfValues() == 16);
  CHECK(rays.HasBuffer("sum_values") && !rays.HasBuffer("nope"));
  CHECK(rays.GetBuffer("attenuationX").GetBufferLength() == 80);
  CHECK(throwsBadValue([&] { rays.GetBuffer("nope"); }, "No channel buffer with requested name"));
  rays.EnableIntersectionData();
  CHECK(rays.NormalX.GetNumberOfValues() == 16);
  rays.Origin.Set(3, vtkm::Vec<float, 3>(1.f, 2.f, 3.f));
  CHECK(rays.OriginY.ReadPortal().Get(3) == 2.f && rays.Origin.Get(3)[2] == 3.f);
  rays.DisableIntersectionData();
  CHECK(rays.NormalX.GetNumberOfValues() == 0);
  RayF copy = rays; // shallow like the reference's array handles
  copy.OriginX.WritePortal().Set(0, 9.f);
  CHECK(rays.OriginX.ReadPortal().Get(0) == 9.f && copy.Origin.Get(0)[0] == 9.f);
}

static void testCameraValidation()
{
  vtkm::rendering::pathtracing::Camera cam;
  CHECK(cam.GetWidth() == 500 && cam.GetHeight() == 500 && cam.seeds.GetNumberOfValues() == 250000);
  // default seeds are the triple-hashed indices (reference Camera.cxx:604-616 == details::WangInit)
  CHECK(cam.seeds.ReadPortal().Get(0) == 413455686u && cam.seeds.ReadPortal().Get(3) == 2223342941u);
  CHECK(throwsBadValue([&] { cam.SetHeight(0); }, "Camera height must be greater than zero."));
  CHECK(throwsBadValue([&] { cam.SetWidth(-3); }, "Camera width must be greater than zero."));
  CHECK(throwsBadValue([&] { cam.SetZoom(0.f); }, "Camera zoom must be greater than zero."));
  CHECK(throwsBadValue([&] { cam.SetFieldOfView(0.f); }, "Camera feild of view must be greater than zero."));
  CHECK(throwsBadValue([&] { cam.SetFieldOfView(181.f); }, "Camera feild of view must be less than 180."));
  cam.ResetIsViewDirty();
  cam.SetPosition(vtkm::Vec<float, 3>(1.f, 2.f, 3.f));
  CHECK(cam.GetIsViewDirty() && cam.GetPosition()[1] == 2.f);
  cam.SetUp(vtkm::Vec<float, 3>(0.f, 2.f, 0.f));
  CHECK(cam.GetUp()[1] == 1.f); // SetUp normalises
  cam.SetWidth(64);
  cam.SetHeight(32);
  CHECK(cam.seeds.GetNumberOfValues() == 64 * 32 && cam.GetFieldOfView() == 30.f);
  CHECK(cam.ToString().find("Width    : 64") != std::string::npos);
  vtkm::Int32 a;
  vtkm::Float32 b;
  CHECK(throwsBadValue([&] { cam.GetPixelData(vtkm::cont::CoordinateSystem(), a, b); }));
}

static void testSceneAndMapperHost()
{
  CornellBox cb;
  cb.buildDataSet();
  CHECK(cb.coord.GetPoints().GetNumberOfValues() == 89 && cb.ds.GetCellSet().GetNumberOfCells() == 23);
  CHECK(cb.matIdx[0].GetNumberOfValues() == 22 && cb.matIdx[1].GetNumberOfValues() == 1 && cb.tex.GetNumberOfValues() == 4);
  CHECK(cb.matType.ReadPortal().Get(3) == 1 && cb.texType.ReadPortal().Get(4) == 0);
  cb.extract();
  CHECK(cb.QuadIds.GetNumberOfValues() == 22 && cb.SphereIds.GetNumberOfValues() == 1 && cb.SphereIds.ReadPortal().Get(0) == 48);
  CHECK(cb.QuadIds.ReadPortal().Get(12)[0] == 13 && cb.QuadIds.ReadPortal().Get(12)[1] == 49);
  vtkm::rendering::MapperPathTracer mapper(10, 5, cb.matIdx, cb.texIdx, cb.matType, cb.texType, cb.tex);
  CHECK(mapper.samplecount == 10 && mapper.depthcount == 5 && mapper.MatIdx == cb.matIdx);
  CHECK(mapper.GetCanvas() == nullptr);
  vtkm::rendering::Canvas plain(8, 8);
  CHECK(throwsBadValue([&] { mapper.SetCanvas(&plain); }, "bad canvas type. Must be CanvasRayTracer"));
  vtkm::rendering::CanvasRayTracer canvas(8, 8);
  mapper.SetCanvas(&canvas);
  CHECK(mapper.GetCanvas() == &canvas && mapper.whichPDF.GetNumberOfValues() == 64);
  auto ex = mapper.extract(cb.ds.GetCellSet());
  CHECK(std::get<3>(ex).GetNumberOfValues() == 22 && std::get<0>(ex).ReadPortal().Get(0) == 48);
  CHECK(std::get<1>(ex).ReadPortal().Get(0) == static_cast<float>(90 / 555.0));
  RayF rays;
  vtkm::cont::ArrayHandle<vtkm::Float32> radii;
  vtkm::cont::ArrayHandle<vtkm::UInt32> seeds;
  CHECK(throwsBadValue([&] { mapper.generateRays(cb.coord, radii, mapper.whichPDF, rays, seeds); }, "fused"));
  vtkm::rendering::pathtracing::PathTracer tracer;
  tracer.AddShapeIntersector(new vtkm::rendering::pathtracing::QuadIntersector());
  CHECK(tracer.GetNumberOfShapes() == 0);
  CHECK(throwsBadValue([&] { tracer.Render(rays); }, "not part of the path-tracing path"));
  tracer.Clear();
  vtkm::rendering::Mapper* copy = mapper.NewCopy();
  CHECK(dynamic_cast<vtkm::rendering::MapperPathTracer*>(copy) != nullptr);
  delete copy;
}

template <typename T>
static void dump(const std::string& path, const T* p, size_t n)
{
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char*>(p), static_cast<std::streamsize>(n * sizeof(T)));
}

static void testGpu(const std::string& prefix)
{
  CornellBox cb;
  cb.buildDataSet();
  const int W = 64, H = 48, spp = 4, depth = 8;
  vtkm::rendering::CanvasRayTracer canvas(W, H);
  vtkm::rendering::Camera cam;
  cam.SetPosition(vec3(278 / 555.0, 278 / 555.0, -800 / 555.0));
  cam.SetFieldOfView(40.);
  cam.SetViewUp(vec3(0, 1, 0));
  cam.SetLookAt(vec3(278 / 555.0, 278 / 555.0, 278 / 555.0));
  vtkm::rendering::MapperPathTracer mapper(spp, depth, cb.matIdx, cb.texIdx, cb.matType, cb.texType, cb.tex);
  mapper.SetCanvas(&canvas);
  vtkm::cont::Field field;
  vtkm::cont::ColorTable ct;
  vtkm::Range sr;
  mapper.RenderCells(cb.ds.GetCellSet(), cb.coord, field, ct, cam, sr);
  CHECK(mapper.GetLastSegments() >= W * H * spp);
  dump(prefix + "_color.bin", reinterpret_cast<const float*>(canvas.GetColorBuffer().GetStorage()), size_t(W) * H * 4);

  // Camera::CreateRays with explicit seeds, then MapperPathTracer::intersect on those rays
  vtkm::rendering::pathtracing::Camera rc;
  rc.SetParameters(cam, canvas);
  for (vtkm::Id i = 0; i < rc.seeds.GetNumberOfValues(); ++i)
    rc.seeds.WritePortal().Set(i, static_cast<unsigned int>(i));
  RayF rays;
  rc.CreateRays(rays, vtkm::Bounds());
  CHECK(rays.NumRays == W * H && rays.HitIdx.ReadPortal().Get(5) == -2 && rays.PixelIdx.ReadPortal().Get(77) == 77);
  CHECK(rays.OriginZ.ReadPortal().Get(0) == static_cast<float>(-800 / 555.0) && rays.MinDistance.ReadPortal().Get(1) == 0.f);
  CHECK(rc.seeds.ReadPortal().Get(0) == 3075307816u); // two wang steps from state 0
  dump(prefix + "_dirx.bin", rays.DirX.GetStorage(), size_t(W) * H);
  for (vtkm::Id i = 0; i < rays.NumRays; ++i)
    rays.Status.WritePortal().Set(i, vtkm::UInt8(1u << 3));
  rays.AddBuffer(depth, "attenuationX"), rays.AddBuffer(depth, "attenuationY"), rays.AddBuffer(depth, "attenuationZ");
  rays.AddBuffer(depth, "emittedX"), rays.AddBuffer(depth, "emittedY"), rays.AddBuffer(depth, "emittedZ");
  vtkm::rendering::raytracing::Vec3View<float> atten{ &rays.GetBuffer("attenuationX").Buffer, &rays.GetBuffer("attenuationY").Buffer,
                                                     &rays.GetBuffer("attenuationZ").Buffer };
  vtkm::rendering::raytracing::Vec3View<float> emit{ &rays.GetBuffer("emittedX").Buffer, &rays.GetBuffer("emittedY").Buffer,
                                                    &rays.GetBuffer("emittedZ").Buffer };
  rays.GetBuffer("attenuationX").InitConst(-7.f);
  mapper.intersect(rays, rays.MinDistance, emit, atten, 2);
  dump(prefix + "_t.bin", rays.Distance.GetStorage(), size_t(W) * H);
  dump(prefix + "_status.bin", rays.Status.GetStorage(), size_t(W) * H);
  int missed = 0;
  for (vtkm::Id i = 0; i < rays.NumRays; ++i)
    if (!(rays.Status.ReadPortal().Get(i) & 8))
    {
      ++missed;
      CHECK(atten.Get(i + rays.NumRays * 2)[0] == 1.f && emit.Get(i + rays.NumRays * 2)[1] == 0.f);
    }
    else
      CHECK(atten.Get(i + rays.NumRays * 2)[0] == -7.f);
  CHECK(missed > 0 && missed < W * H);
  // RenderCellsViews: three cameras in one call, each image the same bits as its own RenderCells call
  {
    std::vector<vtkm::rendering::Camera> cams(3, cam);
    cams[1].SetPosition(vec3(0.9, 0.6, -1.3));
    cams[2].SetPosition(vec3(0.1, 0.4, -1.5));
    cams[2].SetFieldOfView(50.);
    std::vector<vtkm::cont::ArrayHandle<vtkm::Vec<vtkm::Float32, 4>>> colors;
    mapper.RenderCellsViews(cb.ds.GetCellSet(), cb.coord, cams, colors);
    CHECK(colors.size() == 3);
    for (size_t v = 0; v < colors.size(); ++v)
    {
      mapper.RenderCells(cb.ds.GetCellSet(), cb.coord, field, ct, cams[v], sr);
      CHECK(colors[v].GetNumberOfValues() == W * H);
      CHECK(std::memcmp(colors[v].GetStorage(), canvas.GetColorBuffer().GetStorage(), sizeof(float) * 4 * W * H) == 0);
    }
    CHECK(std::memcmp(colors[0].GetStorage(), colors[1].GetStorage(), sizeof(float) * 4 * W * H) != 0);
    // the packed integers equal save()'s arithmetic (main.cc:253-287, 325-384) applied to those sums
    std::vector<unsigned short> pnm;
    mapper.RenderCellsViewsPnm(cb.ds.GetCellSet(), cb.coord, cams, pnm);
    CHECK(pnm.size() == size_t(3) * W * H * 3);
    size_t bad = 0;
    for (size_t v = 0; v < 3; ++v)
      for (vtkm::Id i = 0; i < vtkm::Id(W) * H; ++i)
      {
        auto c = colors[v].ReadPortal().Get(i);
        for (int k = 0; k < 3; ++k)
        {
          float x = (c[k] == c[k]) ? c[k] : 0.f;
          x = std::sqrt(x / static_cast<float>(spp));
          bad += (int(255.99 * x) != int(pnm[(v * size_t(W) * H + size_t(i)) * 3 + k])) ? 1 : 0;
        }
      }
    CHECK(bad == 0);
  }
  // error mapping: invalid canvas size surfaces as ErrorBadValue with the reference's message
  vtkm::rendering::pathtracing::Camera bad;
  CHECK(throwsBadValue([&] { bad.SetWidth(0); }, "Camera width must be greater than zero."));
  b2pt_facade::ReleaseContext();
}

// MapperPathTracer::SetDevices: G GPUs driven from this one process, samples partitioned, sums added by
// b2pt_allreduce (reduce-scatter by k_sum_peers over NVLink peer access + all-gather).  The G-GPU image must equal the
// one-GPU image up to the order of the float additions (identical per-sample radiance, NaN-poisoned pixels included).
static void testMultiGpu(int G)
{
  CornellBox cb;
  cb.buildDataSet();
  const int W = 256, H = 192, spp = 24, depth = 12;
  vtkm::rendering::Camera cam;
  cam.SetPosition(vec3(278 / 555.0, 278 / 555.0, -800 / 555.0));
  cam.SetFieldOfView(40.);
  cam.SetViewUp(vec3(0, 1, 0));
  cam.SetLookAt(vec3(278 / 555.0, 278 / 555.0, 278 / 555.0));
  vtkm::cont::Field field;
  vtkm::cont::ColorTable ct;
  vtkm::Range sr;
  std::vector<float> one, many;
  long long segOne = 0, segMany = 0;
  for (int pass = 0; pass < 2; ++pass)
  {
    vtkm::rendering::CanvasRayTracer canvas(W, H);
    vtkm::rendering::MapperPathTracer mapper(spp, depth, cb.matIdx, cb.texIdx, cb.matType, cb.texType, cb.tex);
    mapper.SetCanvas(&canvas);
    std::vector<int> devs;
    for (int g = 0; g < (pass == 0 ? 1 : G); ++g)
      devs.push_back(g);
    mapper.SetDevices(devs);
    mapper.RenderCells(cb.ds.GetCellSet(), cb.coord, field, ct, cam, sr);
    const float* c = reinterpret_cast<const float*>(canvas.GetColorBuffer().GetStorage());
    (pass == 0 ? one : many).assign(c, c + size_t(W) * H * 4);
    (pass == 0 ? segOne : segMany) = mapper.GetLastSegments();
  }
  CHECK(segOne == segMany && segOne > (long long)W * H * spp);
  size_t bad = 0, nan = 0;
  double maxRel = 0.0;
  for (size_t i = 0; i < one.size(); ++i)
  {
    if ((i & 3) == 3)
      continue; // alpha lane
    const bool n0 = one[i] != one[i], n1 = many[i] != many[i];
    nan += n0;
    if (n0 != n1)
      ++bad;
    else if (!n0)
    {
      const double rel = std::fabs((double)one[i] - (double)many[i]) / std::max(1e-3, std::fabs((double)one[i]));
      maxRel = std::max(maxRel, rel);
      if (rel > 1e-5)
        ++bad;
    }
  }
  std::printf("multi-GPU facade: %d devices, %lld segments, max rel diff %.3g, %zu NaN channels, %zu mismatches\n", G,
              segMany, maxRel, nan, bad);
  CHECK(bad == 0);
  b2pt_facade::ReleaseContext();
}

int main(int argc, char** argv)
{
  if (argc >= 2 && !std::strcmp(argv[1], "--cpu"))
  {
    testChannelBuffer();
    testRay();
    testCameraValidation();
    testSceneAndMapperHost();
  }
  else if (argc >= 3 && !std::strcmp(argv[1], "--gpu"))
  {
    try
    {
      testGpu(argv[2]);
    }
    catch (const vtkm::cont::Error& e)
    {
      std::printf("FAIL exception: %s\n", e.GetMessage().c_str());
      ++g_fail;
    }
  }
  else if (argc >= 3 && !std::strcmp(argv[1], "--multigpu"))
  {
    try
    {
      testMultiGpu(std::atoi(argv[2]));
    }
    catch (const vtkm::cont::Error& e)
    {
      std::printf("FAIL exception: %s\n", e.GetMessage().c_str());
      ++g_fail;
    }
  }
  else
  {
    std::printf("usage: test_facade --cpu | --gpu <prefix> | --multigpu <G>\n");
    return 2;
  }
  std::printf(g_fail ? "%d check(s) FAILED\n" : "all facade checks passed\n", g_fail);
  return g_fail ? 1 : 0;
}
