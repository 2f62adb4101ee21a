// Segmentation over the C ABI (replaces reference segmentation.cpp:879-971).
#include "segmentation.h"
#include "image.h"
#include "defaultContext.h"
#include "../../../include/ssd_gpu.h"
#include <stdexcept>
#include <vector>

namespace stairs
{

namespace
{

// the ABI wants a dense width*height buffer
const uint8_t *dense(const Image &image, std::vector<uint8_t> &tmp)
{
  if(image.step() == image.width())
    return image.ptr();
  tmp.resize(size_t(image.width()) * image.height());
  for(int y = 0; y < image.height(); y++)
    std::copy(image.ptr(0, y), image.ptr(0, y) + image.width(), tmp.begin() + size_t(y) * image.width());
  return tmp.data();
}

} // namespace

Segmentation::FrontEdge Segmentation::detectFrontEdge(const Image &image, const std::string &windowName)
{
  return detectFrontEdge(defaultContext(image.width(), image.height()), image, windowName);
}

Segmentation::Outline Segmentation::detectOutline(const Image &image, int minImgYExtent, double xyRatio, const std::string &windowName)
{
  return detectOutline(defaultContext(image.width(), image.height()), image, minImgYExtent, xyRatio, windowName);
}

Segmentation::FrontEdge Segmentation::detectFrontEdge(ssd_gpu_ctx *ctx, const Image &image, const std::string &)
{
  std::vector<uint8_t> tmp;
  double l[2], r[2];
  int valid = 0;
  if(ssd_gpu_detect_front_edge(ctx, dense(image, tmp), l, r, &valid) != SSD_OK)
    throw std::runtime_error(std::string("ssd_gpu_detect_front_edge: ") + ssd_gpu_last_error(ctx));
  FrontEdge e;
  e.pointLeft = Point2(l[0], l[1]);
  e.pointRight = Point2(r[0], r[1]);
  e.valid = valid != 0;
  return e;
}

Segmentation::Outline Segmentation::detectOutline(ssd_gpu_ctx *ctx, const Image &image, int minImgYExtent, double xyRatio, const std::string &)
{
  std::vector<uint8_t> tmp;
  double q[8];
  int valid = 0;
  if(ssd_gpu_detect_outline(ctx, dense(image, tmp), minImgYExtent, xyRatio, q, &valid) != SSD_OK)
    throw std::runtime_error(std::string("ssd_gpu_detect_outline: ") + ssd_gpu_last_error(ctx));
  Outline o;
  for(int c = 0; c < 4; c++)
    o.quadrilateral[size_t(c)] = Point2(q[c * 2], q[c * 2 + 1]);
  o.valid = valid != 0;
  return o;
}

} // namespace stairs
