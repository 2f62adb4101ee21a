// Host-side value types of the kept class surface (reference types.h:28-115): same names and
// members so callers of the reference's Pointcloud / Transformation / Segmentation / Stairs
// interfaces compile unchanged against this layer.
#pragma once
#include <array>
#include <cstdint>
#include <vector>

namespace stairs
{

using Coordinate_t = double;

template<class T>
struct Point2_
{
  T x{}, y{};
};
using Point2i = Point2_<int>;
using Point2f = Point2_<float>;

struct Point3f
{
  float x{}, y{}, z{};
};

template<int Dim>
struct Point_;

template<>
struct Point_<3>;

template<>
struct Point_<2>
{
  Coordinate_t x = 0, y = 0;
  Point_() = default;
  Point_(Coordinate_t px, Coordinate_t py) : x(px), y(py) {}
  Point_(const Point_<3> &p);
  bool operator!=(const Point_ &o) const { return x != o.x || y != o.y; }
  Coordinate_t &operator[](int i) { return i ? y : x; }
  Coordinate_t operator[](int i) const { return i ? y : x; }
};

template<>
struct Point_<3>
{
  Coordinate_t x = 0, y = 0, z = 0;
  Point_() = default;
  Point_(Coordinate_t px, Coordinate_t py, Coordinate_t pz) : x(px), y(py), z(pz) {}
  Point_(const Point3f &p) : x(p.x), y(p.y), z(p.z) {}
  Point_(const Point_<2> &p, Coordinate_t pz) : x(p.x), y(p.y), z(pz) {}
  Coordinate_t &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  Coordinate_t operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

inline Point_<2>::Point_(const Point_<3> &p) : x(p.x), y(p.y) {}

using Point2 = Point_<2>;
using Point3 = Point_<3>;

struct Size2i
{
  int width = 0, height = 0;
};

template<typename PointType>
using Quadrilateral_ = std::array<PointType, 4>;
using Quadrilateral_t = Quadrilateral_<Point2>;
using Quadrilateralf_t = Quadrilateral_<Point2f>; // projected step corners (reference types.h:115)

} // namespace stairs
