// Test-infrastructure shim (NOT product code): pin-hole project / deproject without distortion.
#ifndef SSD_SHIM_RSUTIL_H
#define SSD_SHIM_RSUTIL_H
#include "hpp/rs_frame.hpp"

inline void rs2_project_point_to_pixel(float pixel[2], const rs2_intrinsics *intrin, const float point[3])
{
  const float x = point[0] / point[2], y = point[1] / point[2];
  pixel[0] = x * intrin->fx + intrin->ppx;
  pixel[1] = y * intrin->fy + intrin->ppy;
}

inline void rs2_deproject_pixel_to_point(float point[3], const rs2_intrinsics *intrin, const float pixel[2], float depth)
{
  const float x = (pixel[0] - intrin->ppx) / intrin->fx;
  const float y = (pixel[1] - intrin->ppy) / intrin->fy;
  point[0] = depth * x;
  point[1] = depth * y;
  point[2] = depth;
}
#endif
