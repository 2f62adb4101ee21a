// Process-wide default GPU contexts for the reference signatures that carry no context:
// Segmentation::detectFrontEdge(const Image&, ...), Segmentation::detectOutline(const Image&, int, double, ...)
// (segmentation.h:35-57) and QuadrilateralTest(const Quadrilateral_t&) (quadrilateralTest.h:35). One context per image
// size, created on first use on device $SSD_GPU_DEVICE (default 0) with the reference's default configuration and the
// identity transformation (neither enters these single-stage calls), destroyed at exit. Not thread-safe beyond creation:
// like every context, calls on it are serialised by the caller.
#pragma once

struct ssd_gpu_ctx;

namespace stairs
{

// throws std::runtime_error when no context can be created (no CUDA device: there is no CPU fallback)
ssd_gpu_ctx *defaultContext(int width, int height);

} // namespace stairs
