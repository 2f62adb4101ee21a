// Test-infrastructure shim (NOT product code): the minimum librealsense2 surface the reference's
// hot-path translation units need to compile by path. Capture is stubbed: a "frame" is a caller
// supplied vertex array. See SURVEY.md Appendix A.
#ifndef SSD_SHIM_RS_FRAME_HPP
#define SSD_SHIM_RS_FRAME_HPP
#include <cstddef>
#include <memory>

struct rs2_intrinsics
{
  int width = 0, height = 0;
  float ppx = 0, ppy = 0, fx = 1, fy = 1;
  int model = 0;
  float coeffs[5] = {0, 0, 0, 0, 0};
};

namespace rs2
{

struct vertex
{
  float x, y, z;
};

struct video_stream_profile
{
  rs2_intrinsics intr;
  rs2_intrinsics get_intrinsics() const { return intr; }
};

struct stream_profile
{
  rs2_intrinsics intr;
  template<class T> T as() const { return T{ intr }; }
};

// One synthetic frame: a W x H grid of vertices owned by the caller.
struct frame_data
{
  int width = 0, height = 0;
  const vertex *vertices = nullptr;
  rs2_intrinsics intr;
};

class frame
{
public:
  frame() {}
  explicit frame(std::shared_ptr<const frame_data> d) : _d(std::move(d)) {}
  const frame_data *shim_data() const { return _d.get(); }
protected:
  std::shared_ptr<const frame_data> _d;
};

class video_frame : public frame
{
public:
  video_frame(const frame &f) : frame(f) {}
  int get_width() const { return _d ? _d->width : 0; }
  int get_height() const { return _d ? _d->height : 0; }
  int get_stride_in_bytes() const { return get_width() * 2; }
  int get_bytes_per_pixel() const { return 2; }
  const void *get_data() const { return nullptr; }
  stream_profile get_profile() const { return stream_profile{ _d ? _d->intr : rs2_intrinsics{} }; }
};

class depth_frame : public video_frame
{
public:
  depth_frame(const frame &f) : video_frame(f) {}
  float get_distance(int, int) const { return 0.f; }
};

class frameset : public frame
{
public:
  frameset(const frame &f) : frame(f) {}
  video_frame get_color_frame() const { return video_frame(*this); }
  video_frame get_infrared_frame() const { return video_frame(*this); }
  depth_frame get_depth_frame() const { return depth_frame(*this); }
};

class points
{
public:
  points() {}
  points(const vertex *v, size_t n) : _v(v), _n(n) {}
  size_t size() const { return _n; }
  const vertex *get_vertices() const { return _v; }
private:
  const vertex *_v = nullptr;
  size_t _n = 0;
};

} // namespace rs2
#endif
