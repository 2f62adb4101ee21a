#include "defaultContext.h"
#include "../../../include/ssd_gpu.h"
#include <cstdlib>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>

namespace stairs
{

namespace
{

struct Registry
{
  std::mutex m;
  std::map<std::pair<int, int>, ssd_gpu_ctx *> ctxs;
  ~Registry()
  {
    for(auto &kv : ctxs)
      ssd_gpu_destroy(kv.second);
  }
};

Registry &registry()
{
  static Registry r;
  return r;
}

} // namespace

ssd_gpu_ctx *defaultContext(int width, int height)
{
  Registry &r = registry();
  std::lock_guard<std::mutex> lock(r.m);
  auto it = r.ctxs.find({ width, height });
  if(it != r.ctxs.end())
    return it->second;
  ssd_gpu_config cfg;
  ssd_gpu_default_config(&cfg, width, height);
  ssd_gpu_transform xf{};
  xf.a[0] = xf.a[4] = xf.a[8] = 1.0;
  xf.ext_a[0] = xf.ext_a[3] = 1.0;
  const char *dev = std::getenv("SSD_GPU_DEVICE");
  ssd_gpu_ctx *ctx = nullptr;
  if(ssd_gpu_create(&cfg, &xf, dev ? std::atoi(dev) : 0, 1, &ctx) != SSD_OK)
    throw std::runtime_error(std::string("stairs::defaultContext: ") + ssd_gpu_last_error(nullptr));
  r.ctxs[{ width, height }] = ctx;
  return ctx;
}

} // namespace stairs
