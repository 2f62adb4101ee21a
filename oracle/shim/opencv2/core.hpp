// Test-infrastructure shim (NOT product code): 8-bit single-channel cv::Mat with shared ownership,
// just enough for the reference's image.h / segmentation.cpp to compile by path.
#ifndef SSD_SHIM_OPENCV_CORE_HPP
#define SSD_SHIM_OPENCV_CORE_HPP
#include <algorithm>
#include <array>
#include <cassert>
#include <vector>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <string>

#define CV_8U 0

namespace cv
{

typedef unsigned char uchar;

class Mat
{
public:
  int rows = 0, cols = 0;
  uchar *data = nullptr;
  size_t step = 0;

  Mat() {}
  Mat(int r, int c, int /*type*/, void *pixels, size_t stp) : rows(r), cols(c), data(static_cast<uchar *>(pixels)), step(stp) {}

  static Mat zeros(int r, int c, int /*type*/)
  {
    Mat m;
    m.rows = r;
    m.cols = c;
    m.step = size_t(c);
    // one guard row: the reference's BEV scatter can write pixel x == width (pointcloud.cpp:81,468)
    const size_t bytes = size_t(r + 1) * size_t(c);
    m._owner = std::shared_ptr<uchar[]>(new uchar[bytes]());
    m.data = m._owner.get();
    return m;
  }

  uchar *ptr(int y = 0) { return data + size_t(y) * step; }
  const uchar *ptr(int y = 0) const { return data + size_t(y) * step; }
  uchar *ptr(int y, int x) { return data + size_t(y) * step + x; }
  const uchar *ptr(int y, int x) const { return data + size_t(y) * step + x; }

private:
  std::shared_ptr<uchar[]> _owner;
};

inline void extractChannel(const Mat &, Mat &, int) {}

} // namespace cv
#endif
