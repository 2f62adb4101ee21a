// Host mirror of the reference's transformation.h:42-126 (Transformation_<Dim>, CameraToWorld,
// WorldToCamera, ToExternalWorld, GeometricTransformation). Built once per camera mounting from
// three calibration points; the 12+7 doubles it holds are what the CUDA kernels consume
// (ssd_gpu_transform). No Boost: the small fixed-size algebra is written out with the same
// operation order as Boost.QVM's generated operators (left-to-right sums of products).
#pragma once
#include "types.h"
#include "../../../include/ssd_gpu.h"

namespace stairs
{

template<int Dim>
using Vector_ = Point_<Dim>;

template<int Dim>
struct Matrix_
{
  Coordinate_t a[Dim][Dim];
  static Matrix_ identity();
  Matrix_ transposed() const;
  Vector_<Dim> column(int j) const;
  void setColumn(int j, const Vector_<Dim> &v);
};

template<int Dim>
using ReferencePoints_ = std::array<Point_<Dim>, Dim>;

template<int Dim>
class Transformation_
{
public:
  using RefPoints = ReferencePoints_<Dim>;
  using Point = Point_<Dim>;
  using Mat = Matrix_<Dim>;
  using Vec = Vector_<Dim>;

  Transformation_() : _a(Mat::identity()), _aInv(_a) {}
  Transformation_(const RefPoints &rp, const RefPoints &rpMapping);
  Transformation_(const RefPoints &triangleInPlane);

  // x -> A*x + b; SrcPointType needs members x,y(,z) (Point3f, Point3, an rs2::vertex look-alike)
  template<typename SrcPointType>
  Point transform(const SrcPointType &x) const;
  Point transformInv(const Point &x) const;

  const Mat &matrix() const { return _a; }
  const Mat &inverseMatrix() const { return _aInv; }
  const Vec &translation() const { return _b; }

private:
  Mat _a, _aInv;
  Vec _b;
};

// boost::qvm::inverse of a 3x3 (adjugate times 1/det), as Transformation_<3>(rp, rpMapping) uses it; throws on det == 0
Matrix_<3> inverseMatrix3(const Matrix_<3> &m);

using Transformation = Transformation_<3>;
using Transformation2D = Transformation_<2>;

template<>
template<typename S>
inline Point3 Transformation_<3>::transform(const S &p) const
{
  const Coordinate_t x = p.x, y = p.y, z = p.z;
  return { ((_a.a[0][0] * x + _a.a[0][1] * y) + _a.a[0][2] * z) + _b.x, ((_a.a[1][0] * x + _a.a[1][1] * y) + _a.a[1][2] * z) + _b.y,
           ((_a.a[2][0] * x + _a.a[2][1] * y) + _a.a[2][2] * z) + _b.z };
}

template<>
template<typename S>
inline Point2 Transformation_<2>::transform(const S &p) const
{
  const Coordinate_t x = p.x, y = p.y;
  return { (_a.a[0][0] * x + _a.a[0][1] * y) + _b.x, (_a.a[1][0] * x + _a.a[1][1] * y) + _b.y };
}

struct CameraToWorld
{
  template<typename SrcPointType>
  Point3 operator()(const SrcPointType &p) const
  {
    return _camera.transform(p);
  }
  const Transformation &_camera;
};

struct WorldToCamera
{
  Point3 operator()(const Point3 &p) const;
  const Transformation &_camera;
};

struct ToExternalWorld
{
  Point3 operator()(const Point3 &p) const;
  const Transformation2D _world;
  const Coordinate_t _worldZ = 0;
};

class GeometricTransformation
{
public:
  using RefPoints = Transformation::RefPoints;

  GeometricTransformation() {}
  GeometricTransformation(const RefPoints &worldPoints, const RefPoints &cameraPoints);
  const CameraToWorld &cameraToWorld() const { return _cameraToWorld; }
  const WorldToCamera &worldToCamera() const { return _worldToCamera; }
  const ToExternalWorld &toExternalWorld() const { return _toExternalWorld; }

  // the doubles the GPU path binds at ssd_gpu_create()
  ssd_gpu_transform abi() const;
  // Transformation_<3>::_aInv of the camera transformation, row-major: what WorldToCamera multiplies by (ssd_gpu_set_overlay)
  void abiInverse(double a_inv[9]) const;

private:
  GeometricTransformation(const GeometricTransformation &) = delete;

  const Transformation _camera;
  const CameraToWorld _cameraToWorld{ _camera };
  const WorldToCamera _worldToCamera{ _camera };
  const ToExternalWorld _toExternalWorld;
};

} // namespace stairs
