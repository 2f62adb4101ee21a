// Stub of the presentation boundary (reference window.h:38-50): the OpenGL window is out of scope, the
// pipeline only ever calls setViewport() on it (pointcloud.cpp:364,384).
#pragma once

namespace stairs
{

namespace viewportId
{
const int grayscale = 0;
const int depth = 1;
const int infrared = 2;
const int numIds = 3;
}

class Window
{
public:
  Window(const char * = "") {}
  void setViewport(int) const {}
  operator bool() const { return true; }
};

} // namespace stairs
