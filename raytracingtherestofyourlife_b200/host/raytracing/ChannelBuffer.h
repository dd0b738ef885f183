// raytracing/ChannelBuffer.h -- facade of vtkm::rendering::raytracing::ChannelBuffer<Precision>
// (reference raytracing/ChannelBuffer.h:54-146, ChannelBuffer.cxx).  Same public surface and error
// behaviour (vtkm::cont::ErrorBadValue with the reference's messages).  On the B200 path these buffers are
// storage only -- the per-depth attenuation/emission layers the reference keeps in them live in registers
// inside the bounce kernel -- so the element-wise utilities below are plain host loops, not kernels.
// Layout: interleaved, value(i, c) = Buffer[i * NumChannels + c] (ChannelBuffer.cxx:176).
#ifndef b2pt_facade_raytracing_ChannelBuffer_h
#define b2pt_facade_raytracing_ChannelBuffer_h

#include <string>

#include <vtkm/cont/ArrayHandle.h>

namespace vtkm
{
namespace rendering
{
namespace raytracing
{

class ChannelBufferOperations;

template <typename Precision>
class ChannelBuffer
{
  friend class ChannelBufferOperations;

protected:
  vtkm::Int32 NumChannels = 4;
  vtkm::Id Size = 0;
  std::string Name = "default";

  static void Require(bool ok, const char* msg)
  {
    if (!ok)
      throw vtkm::cont::ErrorBadValue(msg);
  }

public:
  vtkm::cont::ArrayHandle<Precision> Buffer;

  ChannelBuffer() = default;
  ChannelBuffer(const vtkm::Int32 numChannels, const vtkm::Id size)
  {
    Require(size >= 0, "ChannelBuffer: Size must be greater that -1");
    Require(numChannels >= 0, "ChannelBuffer: NumChannels must be greater that -1");
    NumChannels = numChannels;
    Size = size;
    Buffer.Allocate(Size * NumChannels);
  }

  vtkm::Int32 GetNumChannels() const { return NumChannels; }
  vtkm::Id GetSize() const { return Size; }
  vtkm::Id GetBufferLength() const { return Size * static_cast<vtkm::Id>(NumChannels); }
  void SetName(const std::string name) { Name = name; }
  std::string GetName() const { return Name; }

  void Resize(const vtkm::Id newSize)
  {
    Require(newSize >= 0, "ChannelBuffer resize: Size must be greater than -1");
    Size = newSize;
    Buffer.Allocate(Size * static_cast<vtkm::Id>(NumChannels));
  }
  void SetNumChannels(const vtkm::Int32 numChannels)
  {
    Require(numChannels >= 1, "ChannelBuffer SetNumChannels: numBins must be greater that 0");
    if (NumChannels == numChannels)
      return;
    NumChannels = numChannels;
    Buffer.Allocate(Size * static_cast<vtkm::Id>(NumChannels));
  }

  // element-wise this += other / this *= other
  void AddBuffer(const ChannelBuffer<Precision>& other) { Combine(other, true); }
  void MultiplyBuffer(const ChannelBuffer<Precision>& other) { Combine(other, false); }

  ChannelBuffer<Precision> GetChannel(const vtkm::Int32 channel)
  {
    Require(channel >= 0 && channel < NumChannels, "ChannelBuffer: invalid channel to extract");
    ChannelBuffer<Precision> out(1, Size);
    out.SetName(Name);
    const Precision* src = Buffer.GetStorage();
    Precision* dst = out.Buffer.GetStorage();
    for (vtkm::Id i = 0; i < Size; ++i)
      dst[i] = src[i * NumChannels + channel];
    return out;
  }

  // Scatter the (compacted) entries back to a buffer of outputSize entries; untouched entries take the
  // per-channel signature / the constant initValue.
  ChannelBuffer<Precision> ExpandBuffer(vtkm::cont::ArrayHandle<vtkm::Id> sparseIndexes, const vtkm::Id outputSize,
                                        vtkm::cont::ArrayHandle<Precision> signature)
  {
    Require(signature.GetNumberOfValues() == NumChannels,
            "ChannelBuffer: number of bins in sourse signature must match NumChannels");
    ChannelBuffer<Precision> out(NumChannels, outputSize);
    out.SetName(Name);
    out.InitChannels(signature);
    ScatterInto(out, sparseIndexes);
    return out;
  }
  ChannelBuffer<Precision> ExpandBuffer(vtkm::cont::ArrayHandle<vtkm::Id> sparseIndexes, const vtkm::Id outputSize,
                                        Precision initValue = 1.f)
  {
    ChannelBuffer<Precision> out(NumChannels, outputSize);
    out.SetName(Name);
    out.InitConst(initValue);
    ScatterInto(out, sparseIndexes);
    return out;
  }

  ChannelBuffer<Precision> Copy()
  {
    ChannelBuffer<Precision> out(NumChannels, Size);
    out.SetName(Name);
    const Precision* src = Buffer.GetStorage();
    Precision* dst = out.Buffer.GetStorage();
    for (vtkm::Id i = 0; i < GetBufferLength(); ++i)
      dst[i] = src[i];
    return out;
  }

  void InitConst(const Precision value)
  {
    Precision* p = Buffer.GetStorage();
    for (vtkm::Id i = 0; i < GetBufferLength(); ++i)
      p[i] = value;
  }
  void InitChannels(const vtkm::cont::ArrayHandle<Precision>& signature)
  {
    Require(signature.GetNumberOfValues() == NumChannels,
            "ChannelBuffer: number of bins in sourse signature must match NumChannels");
    Precision* p = Buffer.GetStorage();
    const Precision* s = signature.GetStorage();
    for (vtkm::Id i = 0; i < GetBufferLength(); ++i)
      p[i] = s[i % NumChannels];
  }

  // (v - min) / (max - min) over the whole buffer, optionally inverted; a constant buffer is scaled by its
  // own value like the reference (ChannelBuffer.cxx:351-362).
  void Normalize(bool invert)
  {
    const vtkm::Id n = GetBufferLength();
    if (n == 0)
      return;
    Precision* p = Buffer.GetStorage();
    Precision lo = p[0], hi = p[0];
    for (vtkm::Id i = 1; i < n; ++i)
    {
      lo = p[i] < lo ? p[i] : lo;
      hi = p[i] > hi ? p[i] : hi;
    }
    const Precision scale = (hi - lo == 0.) ? lo : Precision(1.f / (hi - lo));
    for (vtkm::Id i = 0; i < n; ++i)
    {
      Precision v = (p[i] - lo) * scale;
      p[i] = invert ? Precision(1.f - v) : v;
    }
  }

private:
  void Combine(const ChannelBuffer<Precision>& other, bool add)
  {
    Require(NumChannels == other.NumChannels, "ChannelBuffer add: number of channels must be equal");
    Require(Size == other.Size, "ChannelBuffer add: size must be equal");
    Precision* a = Buffer.GetStorage();
    const Precision* b = other.Buffer.GetStorage();
    for (vtkm::Id i = 0; i < GetBufferLength(); ++i)
      a[i] = add ? a[i] + b[i] : a[i] * b[i];
  }
  void ScatterInto(ChannelBuffer<Precision>& out, const vtkm::cont::ArrayHandle<vtkm::Id>& sparse) const
  {
    const Precision* src = Buffer.GetStorage();
    Precision* dst = out.Buffer.GetStorage();
    const vtkm::Id* idx = sparse.GetStorage();
    for (vtkm::Id i = 0; i < Size; ++i)
    {
      Require(i < sparse.GetNumberOfValues() && idx[i] >= 0 && idx[i] < out.Size,
              "ChannelBuffer expand: sparse index out of range");
      for (vtkm::Int32 c = 0; c < NumChannels; ++c)
        dst[idx[i] * NumChannels + c] = src[i * NumChannels + c];
    }
  }
};

} // namespace raytracing
} // namespace rendering
} // namespace vtkm
#endif
