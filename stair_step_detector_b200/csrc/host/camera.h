// Stub of the capture boundary (reference camera.h:31-86). The RealSense pipeline is out of scope
// (BASELINE.json north_star: capture stubbed); a DepthFrame here is a view of W x H packed {x,y,z} float
// vertices -- what rs2::pointcloud::calculate hands to the reference at pointcloud.cpp:138 -- in host or
// device memory.
#pragma once
#include "types.h"
#include <cstddef>

namespace stairs
{

class Camera
{
public:
  struct DepthFrame
  {
    const float *vertices = nullptr; // width*height*3 floats, row-major pixel order, invalid pixel = (0,0,0)
    int w = 0, h = 0;
    bool onDevice = false;

    DepthFrame() = default;
    DepthFrame(const float *xyz, int width, int height, bool deviceMemory = false) : vertices(xyz), w(width), h(height), onDevice(deviceMemory) {}
    int width() const { return w; }
    int height() const { return h; }
    size_t size() const { return size_t(w) * size_t(h); }
  };
};

} // namespace stairs
