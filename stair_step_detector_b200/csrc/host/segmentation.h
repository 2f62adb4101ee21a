// Host mirror of the reference's 2-D outline detector interface (segmentation.h:32-58): same static
// functions and result structs. The work runs in the k_outline / front-edge CUDA code on the plateau's
// top-down image. The reference's signatures (no context argument) run on a process-wide default context for the
// image's size (defaultContext.h); the overloads with a leading ssd_gpu_ctx* use the caller's.
#pragma once
#include "types.h"
#include <string>

struct ssd_gpu_ctx;

namespace stairs
{

class Image;

class Segmentation
{
public:
  struct FrontEdge
  {
    Point2 pointLeft, pointRight;
    bool valid = false;
  };
  static FrontEdge detectFrontEdge(const Image &image, const std::string &windowName = ""); // segmentation.h:40
  static FrontEdge detectFrontEdge(ssd_gpu_ctx *ctx, const Image &image, const std::string &windowName = "");

  struct Outline
  {
    // vertex order: 0 front-left, 1 front-right, 2 back-left, 3 back-right (image y grows towards the front)
    Quadrilateral_t quadrilateral;
    bool valid = false;
  };
  static Outline detectOutline(const Image &image, int minImgYExtent, double xyRatio, const std::string &windowName = ""); // segmentation.h:57
  static Outline detectOutline(ssd_gpu_ctx *ctx, const Image &image, int minImgYExtent, double xyRatio, const std::string &windowName = "");
};

} // namespace stairs
