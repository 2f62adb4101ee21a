// Host mirror of the reference's transformation.cpp (rigid transforms from calibration points):
//   Transformation_<3>(triangleInPlane)   reference transformation.cpp:108-157
//   Transformation_<2>(rp, rpMapping)     reference transformation.cpp:94-106 with makeRotation :65-90
//   Transformation_<3>(rp, rpMapping)     reference transformation.cpp:159-183 (affine, debugging only)
//   GeometricTransformation ctor          reference transformation.cpp:196-215
#include "transformation.h"
#include <cmath>
#include <stdexcept>

namespace stairs
{

namespace
{

using V3 = Vector_<3>;
using V2 = Vector_<2>;

inline V3 sub(const V3 &a, const V3 &b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V2 sub(const V2 &a, const V2 &b) { return { a.x - b.x, a.y - b.y }; }
inline V3 neg(const V3 &a) { return { -a.x, -a.y, -a.z }; }
inline Coordinate_t dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(const V3 &a, const V3 &b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }

// QVM normalized(): scale by the reciprocal magnitude
inline V3 unit(const V3 &a)
{
  const Coordinate_t m2 = a.x * a.x + a.y * a.y + a.z * a.z;
  if(m2 == 0)
    throw std::invalid_argument("zero-length vector");
  const Coordinate_t rm = 1 / std::sqrt(m2);
  return { a.x * rm, a.y * rm, a.z * rm };
}
inline V2 unit(const V2 &a)
{
  const Coordinate_t m2 = a.x * a.x + a.y * a.y;
  if(m2 == 0)
    throw std::invalid_argument("zero-length vector");
  const Coordinate_t rm = 1 / std::sqrt(m2);
  return { a.x * rm, a.y * rm };
}

Coordinate_t det3(const Matrix_<3> &m)
{
  return m.a[0][0] * (m.a[1][1] * m.a[2][2] - m.a[1][2] * m.a[2][1]) - m.a[0][1] * (m.a[1][0] * m.a[2][2] - m.a[1][2] * m.a[2][0]) +
         m.a[0][2] * (m.a[1][0] * m.a[2][1] - m.a[1][1] * m.a[2][0]);
}

Matrix_<3> inverse3(const Matrix_<3> &m)
{
  const Coordinate_t det = det3(m);
  if(det == 0)
    throw std::invalid_argument("singular matrix");
  const Coordinate_t f = 1 / det;
  Matrix_<3> r;
  r.a[0][0] = f * (m.a[1][1] * m.a[2][2] - m.a[1][2] * m.a[2][1]);
  r.a[0][1] = f * (m.a[0][2] * m.a[2][1] - m.a[0][1] * m.a[2][2]);
  r.a[0][2] = f * (m.a[0][1] * m.a[1][2] - m.a[0][2] * m.a[1][1]);
  r.a[1][0] = f * (m.a[1][2] * m.a[2][0] - m.a[1][0] * m.a[2][2]);
  r.a[1][1] = f * (m.a[0][0] * m.a[2][2] - m.a[0][2] * m.a[2][0]);
  r.a[1][2] = f * (m.a[0][2] * m.a[1][0] - m.a[0][0] * m.a[1][2]);
  r.a[2][0] = f * (m.a[1][0] * m.a[2][1] - m.a[1][1] * m.a[2][0]);
  r.a[2][1] = f * (m.a[0][1] * m.a[2][0] - m.a[0][0] * m.a[2][1]);
  r.a[2][2] = f * (m.a[0][0] * m.a[1][1] - m.a[0][1] * m.a[1][0]);
  return r;
}

Matrix_<3> mul3(const Matrix_<3> &a, const Matrix_<3> &b)
{
  Matrix_<3> r;
  for(int i = 0; i < 3; i++)
    for(int j = 0; j < 3; j++)
      r.a[i][j] = a.a[i][0] * b.a[0][j] + a.a[i][1] * b.a[1][j] + a.a[i][2] * b.a[2][j];
  return r;
}

} // namespace

template<int Dim>
Matrix_<Dim> Matrix_<Dim>::identity()
{
  Matrix_ m;
  for(int i = 0; i < Dim; i++)
    for(int j = 0; j < Dim; j++)
      m.a[i][j] = i == j ? 1 : 0;
  return m;
}

template<int Dim>
Matrix_<Dim> Matrix_<Dim>::transposed() const
{
  Matrix_ m;
  for(int i = 0; i < Dim; i++)
    for(int j = 0; j < Dim; j++)
      m.a[i][j] = a[j][i];
  return m;
}

template<int Dim>
Vector_<Dim> Matrix_<Dim>::column(int j) const
{
  Vector_<Dim> v;
  for(int i = 0; i < Dim; i++)
    v[i] = a[i][j];
  return v;
}

template<int Dim>
void Matrix_<Dim>::setColumn(int j, const Vector_<Dim> &v)
{
  for(int i = 0; i < Dim; i++)
    a[i][j] = v[i];
}

template struct Matrix_<2>;
template struct Matrix_<3>;

// Plane through three camera-space points -> camera-dependent world frame: z up (towards the camera),
// y = in-plane direction with no camera-x component, origin at the camera's foot point.
template<>
Transformation_<3>::Transformation_(const RefPoints &tri)
{
  const Vec p = tri[0];
  const Vec n0 = unit(cross(sub(tri[1], p), sub(tri[2], p)));
  const Vec zBase = neg(n0);
  const Vec yBase = unit(Vec{ 0, -zBase.z / zBase.y, 1 });
  const Vec xBase = cross(yBase, zBase);

  _aInv.setColumn(0, xBase);
  _aInv.setColumn(1, yBase);
  _aInv.setColumn(2, zBase);
  _a = _aInv.transposed(); // rotation: inverse = transpose

  const Coordinate_t distFromOrigin = dot(p, n0);
  if(!(distFromOrigin > 0))
    throw std::invalid_argument("calibration plane is behind the camera");
  _b = Vec{ 0, 0, distFromOrigin };
}

// 2-D rigid map taking rpMapping[0..1] onto rp[0..1] (rotation from the two direction vectors)
template<>
Transformation_<2>::Transformation_(const RefPoints &rp, const RefPoints &rpMapping)
{
  const Vec d = unit(sub(rp[1], rp[0]));
  const Vec dm = unit(sub(rpMapping[1], rpMapping[0]));
  const Coordinate_t c = d.x * dm.x + d.y * dm.y;
  const Coordinate_t s = d.y * dm.x - d.x * dm.y;
  _a.setColumn(0, Vec{ c, s });
  _a.setColumn(1, Vec{ -s, c });
  _aInv = _a.transposed();
  const Vec &m0 = rpMapping.front();
  _b = sub(rp.front(), Vec{ _a.a[0][0] * m0.x + _a.a[0][1] * m0.y, _a.a[1][0] * m0.x + _a.a[1][1] * m0.y });
}

// General affine map from two point triples (the reference keeps one "for debugging and testing")
template<>
Transformation_<3>::Transformation_(const RefPoints &rp, const RefPoints &rpMapping)
{
  auto linear = [](const RefPoints &t)
  {
    const Vec u = sub(t[1], t[0]), v = sub(t[2], t[0]);
    Mat m;
    m.setColumn(0, u);
    m.setColumn(1, v);
    m.setColumn(2, cross(u, v));
    return m;
  };
  _a = mul3(linear(rp), inverse3(linear(rpMapping)));
  _aInv = inverse3(_a);
  const Vec &m0 = rpMapping.front();
  _b = sub(rp.front(), Vec{ _a.a[0][0] * m0.x + _a.a[0][1] * m0.y + _a.a[0][2] * m0.z, _a.a[1][0] * m0.x + _a.a[1][1] * m0.y + _a.a[1][2] * m0.z,
                            _a.a[2][0] * m0.x + _a.a[2][1] * m0.y + _a.a[2][2] * m0.z });
}

template<>
Point3 Transformation_<3>::transformInv(const Point &x) const
{
  const Vec d = sub(x, _b);
  return { _aInv.a[0][0] * d.x + _aInv.a[0][1] * d.y + _aInv.a[0][2] * d.z, _aInv.a[1][0] * d.x + _aInv.a[1][1] * d.y + _aInv.a[1][2] * d.z,
           _aInv.a[2][0] * d.x + _aInv.a[2][1] * d.y + _aInv.a[2][2] * d.z };
}

template<>
Point2 Transformation_<2>::transformInv(const Point &x) const
{
  const Vec d = sub(x, _b);
  return { _aInv.a[0][0] * d.x + _aInv.a[0][1] * d.y, _aInv.a[1][0] * d.x + _aInv.a[1][1] * d.y };
}

Point3 WorldToCamera::operator()(const Point3 &p) const
{
  return _camera.transformInv(p);
}

Point3 ToExternalWorld::operator()(const Point3 &p) const
{
  return { _world.transform(Point2(p)), _worldZ + p.z };
}

GeometricTransformation::GeometricTransformation(const RefPoints &worldPoints, const RefPoints &cameraPoints)
: _camera(cameraPoints),
  _toExternalWorld{ Transformation2D({ Point2(worldPoints[0]), Point2(worldPoints[1]) },
                                     { Point2(_cameraToWorld(cameraPoints[0])), Point2(_cameraToWorld(cameraPoints[1])) }),
                    worldPoints[0].z }
{
}

ssd_gpu_transform GeometricTransformation::abi() const
{
  ssd_gpu_transform t{};
  for(int i = 0; i < 3; i++)
  {
    for(int j = 0; j < 3; j++)
      t.a[i * 3 + j] = _camera.matrix().a[i][j];
    t.b[i] = _camera.translation()[i];
  }
  for(int i = 0; i < 2; i++)
  {
    for(int j = 0; j < 2; j++)
      t.ext_a[i * 2 + j] = _toExternalWorld._world.matrix().a[i][j];
    t.ext_b[i] = _toExternalWorld._world.translation()[i];
  }
  t.ext_z = _toExternalWorld._worldZ;
  return t;
}

Matrix_<3> inverseMatrix3(const Matrix_<3> &m)
{
  return inverse3(m);
}

void GeometricTransformation::abiInverse(double a_inv[9]) const
{
  for(int i = 0; i < 3; i++)
    for(int j = 0; j < 3; j++)
      a_inv[i * 3 + j] = _camera.inverseMatrix().a[i][j];
}

} // namespace stairs
