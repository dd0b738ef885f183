#include "b2pt_facade.h"

#include <cstdlib>
#include <mutex>

namespace b2pt_facade
{
namespace
{
b2pt_ctx* g_ctx = nullptr;
std::mutex g_mutex;
}

b2pt_ctx* Context()
{
  std::lock_guard<std::mutex> lock(g_mutex);
  if (!g_ctx)
  {
    const char* dev = std::getenv("B2PT_DEVICE");
    int err = 0;
    g_ctx = b2pt_create(dev ? std::atoi(dev) : 0, &err);
    if (!g_ctx)
      Check(err);
  }
  return g_ctx;
}

void ReleaseContext()
{
  std::lock_guard<std::mutex> lock(g_mutex);
  if (g_ctx)
    b2pt_destroy(g_ctx);
  g_ctx = nullptr;
}
} // namespace b2pt_facade
