// raytracing/ChannelBufferOperations.h -- facade of ChannelBufferOperations
// (reference raytracing/ChannelBufferOperations.h:103-148): Compact / InitChannels / InitConst.
// Compact is the reference's scan + scatter stream compaction; the model for the device-side
// ballot/prefix-sum ray-queue compaction inside k_bounce, kept here on host arrays for API parity.
#ifndef b2pt_facade_raytracing_ChannelBufferOperations_h
#define b2pt_facade_raytracing_ChannelBufferOperations_h

#include "ChannelBuffer.h"

namespace vtkm
{
namespace rendering
{
namespace raytracing
{

class ChannelBufferOperations
{
public:
  // Keep entry i iff masks[i] != 0, preserving order; newSize must equal the number of kept entries.
  template <typename Precision>
  static void Compact(ChannelBuffer<Precision>& buffer, vtkm::cont::ArrayHandle<vtkm::UInt8>& masks,
                      const vtkm::Id& newSize)
  {
    if (masks.GetNumberOfValues() != buffer.Size)
      throw vtkm::cont::ErrorBadValue("ChannelBuffer compact: mask size must equal buffer size");
    vtkm::cont::ArrayHandle<Precision> packed;
    packed.Allocate(newSize * buffer.NumChannels);
    const Precision* src = buffer.Buffer.GetStorage();
    Precision* dst = packed.GetStorage();
    const vtkm::UInt8* m = masks.GetStorage();
    vtkm::Id out = 0;
    for (vtkm::Id i = 0; i < buffer.Size; ++i)
    {
      if (!m[i])
        continue;
      if (out >= newSize)
        throw vtkm::cont::ErrorBadValue("ChannelBuffer compact: newSize smaller than the number of kept entries");
      for (vtkm::Int32 c = 0; c < buffer.NumChannels; ++c)
        dst[out * buffer.NumChannels + c] = src[i * buffer.NumChannels + c];
      ++out;
    }
    buffer.Buffer = packed;
    buffer.Size = newSize;
  }

  template <typename Device, typename Precision>
  static void InitChannels(ChannelBuffer<Precision>& buffer, vtkm::cont::ArrayHandle<Precision> sourceSignature,
                           Device)
  {
    buffer.InitChannels(sourceSignature);
  }

  template <typename Device, typename Precision>
  static void InitConst(ChannelBuffer<Precision>& buffer, const Precision value, Device)
  {
    buffer.InitConst(value);
  }
};

} // namespace raytracing
} // namespace rendering
} // namespace vtkm
#endif
