// QuadrilateralTest over the C ABI (replaces reference quadrilateralTest.cpp:275-451).
#include "quadrilateralTest.h"
#include "defaultContext.h"
#include "../../../include/ssd_gpu.h"
#include <stdexcept>
#include <string>

namespace stairs
{

QuadrilateralTest::QuadrilateralTest(const Quadrilateral_t &q) : QuadrilateralTest(defaultContext(640, 480), q) {}

QuadrilateralTest::QuadrilateralTest(ssd_gpu_ctx *ctx, const Quadrilateral_t &q) : _ctx(ctx)
{
  for(int c = 0; c < 4; c++)
  {
    _quad[c * 2] = q[size_t(c)].x;
    _quad[c * 2 + 1] = q[size_t(c)].y;
  }
  // constructing evaluates the segment map once so that degenerate input is reported here, as in the reference
  const double probe[2] = { q[0].x, q[0].y };
  uint8_t inside = 0;
  int status = 0;
  if(ssd_gpu_points_in_quad(_ctx, _quad, probe, 1, &inside, &status) != SSD_OK)
    throw std::runtime_error(std::string("ssd_gpu_points_in_quad: ") + ssd_gpu_last_error(_ctx));
  if(status)
    throw std::invalid_argument("Quadrilateral is not convex or has a degenerate segment map.");
}

std::vector<uint8_t> QuadrilateralTest::arePointsWithin(const std::vector<Point2> &points) const
{
  std::vector<uint8_t> inside(points.size());
  if(points.empty())
    return inside;
  static_assert(sizeof(Point2) == 2 * sizeof(double), "Point2 must be two packed doubles");
  int status = 0;
  if(ssd_gpu_points_in_quad(_ctx, _quad, reinterpret_cast<const double *>(points.data()), int(points.size()), inside.data(), &status) != SSD_OK)
    throw std::runtime_error(std::string("ssd_gpu_points_in_quad: ") + ssd_gpu_last_error(_ctx));
  return inside;
}

bool QuadrilateralTest::isPointWithin(const Point2 &point) const
{
  return arePointsWithin({ point }).front() != 0;
}

} // namespace stairs
