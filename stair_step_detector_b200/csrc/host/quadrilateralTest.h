// Host mirror of the reference's point-in-convex-quadrilateral test (quadrilateralTest.h:32-36). The
// segment map and the per-cell half-plane tests are built and evaluated on the GPU (quadtest_init /
// quadtest_within in ssd_device.cuh); this class batches points through ssd_gpu_points_in_quad.
// Like the reference's constructor it throws std::invalid_argument for quadrilaterals that are not convex,
// have no extent, or produce an inconsistent segment map.
#pragma once
#include "types.h"
#include <cstdint>
#include <vector>

struct ssd_gpu_ctx;

namespace stairs
{

class QuadrilateralTest
{
public:
  explicit QuadrilateralTest(const Quadrilateral_t &q); // quadrilateralTest.h:35 (default context, defaultContext.h)
  QuadrilateralTest(ssd_gpu_ctx *ctx, const Quadrilateral_t &q);
  bool isPointWithin(const Point2 &point) const;
  std::vector<uint8_t> arePointsWithin(const std::vector<Point2> &points) const;

private:
  ssd_gpu_ctx *_ctx;
  double _quad[8];
};

} // namespace stairs
