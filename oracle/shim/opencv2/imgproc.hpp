// Test-infrastructure shim (NOT product code): cv::morphologyEx(MORPH_CLOSE) with the default
// 3x3 rectangular element, centre anchor, one iteration, default border
// (dilate treats outside as 0, erode treats outside as 255). Checked bit-equal to
// cv2.morphologyEx(img, cv2.MORPH_CLOSE, None) of opencv-python 4.13 in tests/test_oracle_close.py.
#ifndef SSD_SHIM_OPENCV_IMGPROC_HPP
#define SSD_SHIM_OPENCV_IMGPROC_HPP
#include "core.hpp"
#include <vector>

namespace cv
{

enum { MORPH_ERODE = 0, MORPH_DILATE = 1, MORPH_OPEN = 2, MORPH_CLOSE = 3 };

// 3x3 box max/min is separable: 1x3 along x, then 3x1 along y (identical result).
template<class Op>
inline void shim_box3(const uchar *src, size_t sstep, uchar *dst, size_t dstep, int W, int H, Op op)
{
  std::vector<uchar> tmp(size_t(W) * H);
  for(int y = 0; y < H; y++)
  {
    const uchar *s = src + size_t(y) * sstep;
    uchar *t = tmp.data() + size_t(y) * W;
    if(W == 1)
    {
      t[0] = s[0];
      continue;
    }
    t[0] = op(s[0], s[1]);
    for(int x = 1; x < W - 1; x++)
      t[x] = op(op(s[x - 1], s[x]), s[x + 1]);
    t[W - 1] = op(s[W - 2], s[W - 1]);
  }
  for(int y = 0; y < H; y++)
  {
    const uchar *t0 = tmp.data() + size_t(y > 0 ? y - 1 : y) * W;
    const uchar *t1 = tmp.data() + size_t(y) * W;
    const uchar *t2 = tmp.data() + size_t(y < H - 1 ? y + 1 : y) * W;
    uchar *d = dst + size_t(y) * dstep;
    for(int x = 0; x < W; x++)
      d[x] = op(op(t0[x], t1[x]), t2[x]);
  }
}

inline void morphologyEx(const Mat &src, const Mat &dst, int op, const Mat & /*kernel*/)
{
  assert(op == MORPH_CLOSE);
  (void)op;
  const int W = src.cols, H = src.rows;
  std::vector<uchar> dil(size_t(W) * H);
  // ignoring out-of-image neighbours == border 0 for dilate and border 255 for erode
  shim_box3(src.data, src.step, dil.data(), size_t(W), W, H, [](uchar a, uchar b) { return a > b ? a : b; });
  Mat &out = const_cast<Mat &>(dst);
  shim_box3(dil.data(), size_t(W), out.data, out.step, W, H, [](uchar a, uchar b) { return a < b ? a : b; });
}

} // namespace cv
#endif
