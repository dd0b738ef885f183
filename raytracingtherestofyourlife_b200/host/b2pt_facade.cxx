#include "b2pt_facade.h"

#include <cstdlib>
#include <map>
#include <mutex>

namespace b2pt_facade
{
namespace
{
std::map<int, b2pt_ctx*> g_ctx; // one libb2pt context per GPU, created on first use
std::mutex g_mutex;
}

int DefaultDevice()
{
  const char* dev = std::getenv("B2PT_DEVICE");
  return dev ? std::atoi(dev) : 0;
}

b2pt_ctx* Context(int device)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  if (device < 0)
    device = DefaultDevice();
  b2pt_ctx*& ctx = g_ctx[device];
  if (!ctx)
  {
    int err = 0;
    ctx = b2pt_create(device, &err);
    if (!ctx)
    {
      g_ctx.erase(device);
      Check(err);
    }
  }
  return ctx;
}

void ReleaseContext()
{
  std::lock_guard<std::mutex> lock(g_mutex);
  for (auto& kv : g_ctx)
    if (kv.second)
      b2pt_destroy(kv.second);
  g_ctx.clear();
}
} // namespace b2pt_facade
