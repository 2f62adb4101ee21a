// Test-infrastructure shim (NOT product code). See hpp/rs_frame.hpp.
#ifndef SSD_SHIM_RS_HPP
#define SSD_SHIM_RS_HPP
#include "hpp/rs_frame.hpp"

namespace rs2
{

// SDK deprojection is outside the hot path: the synthetic frame already carries vertices.
class pointcloud
{
public:
  pointcloud() {} // user-provided: it is a const member of stairs::Pointcloud
  points calculate(const frame &f) const
  {
    const frame_data *d = f.shim_data();
    return d ? points(d->vertices, size_t(d->width) * size_t(d->height)) : points();
  }
};

class colorizer
{
public:
  colorizer() {}
  explicit colorizer(int) {}
  video_frame colorize(const frame &f) const { return video_frame(f); }
};

class pipeline
{
public:
  frameset wait_for_frames() { return frameset(frame()); }
};

} // namespace rs2
#endif
