// 8-bit single-channel image, the type Segmentation consumes (reference image.h:30-84 wraps cv::Mat;
// OpenCV is not part of this path, so this owns a plain buffer).
#pragma once
#include "types.h"
#include <cstdint>
#include <vector>

namespace stairs
{

class Image
{
public:
  Image() {}
  Image(const Size2i &s) : Image(s.width, s.height) {}
  Image(int width, int height) : _w(width), _h(height), _store(size_t(width) * height, 0), _p(_store.data()), _step(width) {}
  Image(int width, int height, void *pixels, int step) : _w(width), _h(height), _p(static_cast<uint8_t *>(pixels)), _step(step) {}

  int width() const { return _w; }
  int height() const { return _h; }
  uint8_t *pixels() const { return _p; }
  int step() const { return _step; }
  const uint8_t *ptr() const { return _p; }
  const uint8_t *ptr(int x, int y) const { return _p + size_t(y) * _step + x; }
  uint8_t *ptr(int x, int y) { return _p + size_t(y) * _step + x; }
  uint8_t *ptr(const Point2i &p) { return ptr(p.x, p.y); }

private:
  int _w = 0, _h = 0;
  std::vector<uint8_t> _store;
  uint8_t *_p = nullptr;
  int _step = 0;
};

} // namespace stairs
