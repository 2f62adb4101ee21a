// Test-infrastructure shim (NOT product code): the slice of Boost.QVM (header-only, absent from this
// image, version unpinned by the reference) that transformation.h/.cpp, segmentation.cpp and
// qvmTraits.h use. Arithmetic follows QVM's generated operators: fixed left-to-right sums of
// products, element type = common type of the operands (float vertex * double matrix -> double).
#ifndef SSD_SHIM_BOOST_QVM_ALL_HPP
#define SSD_SHIM_BOOST_QVM_ALL_HPP
#include <cassert>
#include <cmath>
#include <limits>
#include <stdexcept>
#include <type_traits>

namespace boost { namespace qvm
{

// ---- traits ----
template<class V>
struct vec_traits
{
  static int const dim = 0;
  typedef void scalar_type;
};

template<class V, class S, int D>
struct vec_traits_defaults
{
  typedef S scalar_type;
  static int const dim = D;
  template<int I>
  static S read_element(V const &v)
  {
    return vec_traits<V>::template write_element<I>(const_cast<V &>(v));
  }
};

template<class A, class B, int D>
struct deduce_vec2
{
  typedef A type;
};

template<class T, int R, int C>
struct mat
{
  T a[R][C];
};

template<class V>
concept Vec2 = (vec_traits<V>::dim == 2);
template<class V>
concept Vec3 = (vec_traits<V>::dim == 3);

template<int I, class V>
inline typename vec_traits<V>::scalar_type rd(V const &v)
{
  return vec_traits<V>::template read_element<I>(v);
}
template<int I, class V>
inline typename vec_traits<V>::scalar_type &wr(V &v)
{
  return vec_traits<V>::template write_element<I>(v);
}

// ---- vec (+,-) vec ----
template<Vec2 A, Vec2 B>
inline typename deduce_vec2<A, B, 2>::type operator+(A const &a, B const &b)
{
  typename deduce_vec2<A, B, 2>::type r;
  wr<0>(r) = rd<0>(a) + rd<0>(b);
  wr<1>(r) = rd<1>(a) + rd<1>(b);
  return r;
}
template<Vec3 A, Vec3 B>
inline typename deduce_vec2<A, B, 3>::type operator+(A const &a, B const &b)
{
  typename deduce_vec2<A, B, 3>::type r;
  wr<0>(r) = rd<0>(a) + rd<0>(b);
  wr<1>(r) = rd<1>(a) + rd<1>(b);
  wr<2>(r) = rd<2>(a) + rd<2>(b);
  return r;
}
template<Vec2 A, Vec2 B>
inline typename deduce_vec2<A, B, 2>::type operator-(A const &a, B const &b)
{
  typename deduce_vec2<A, B, 2>::type r;
  wr<0>(r) = rd<0>(a) - rd<0>(b);
  wr<1>(r) = rd<1>(a) - rd<1>(b);
  return r;
}
template<Vec3 A, Vec3 B>
inline typename deduce_vec2<A, B, 3>::type operator-(A const &a, B const &b)
{
  typename deduce_vec2<A, B, 3>::type r;
  wr<0>(r) = rd<0>(a) - rd<0>(b);
  wr<1>(r) = rd<1>(a) - rd<1>(b);
  wr<2>(r) = rd<2>(a) - rd<2>(b);
  return r;
}
template<Vec3 A>
inline A operator-(A const &a)
{
  A r;
  wr<0>(r) = -rd<0>(a);
  wr<1>(r) = -rd<1>(a);
  wr<2>(r) = -rd<2>(a);
  return r;
}
template<Vec2 A>
inline A operator-(A const &a)
{
  A r;
  wr<0>(r) = -rd<0>(a);
  wr<1>(r) = -rd<1>(a);
  return r;
}
template<Vec2 A, Vec2 B>
inline A &operator+=(A &a, B const &b)
{
  wr<0>(a) += rd<0>(b);
  wr<1>(a) += rd<1>(b);
  return a;
}
template<Vec3 A, Vec3 B>
inline A &operator+=(A &a, B const &b)
{
  wr<0>(a) += rd<0>(b);
  wr<1>(a) += rd<1>(b);
  wr<2>(a) += rd<2>(b);
  return a;
}
template<Vec2 A>
inline A &operator/=(A &a, typename vec_traits<A>::scalar_type s)
{
  wr<0>(a) /= s;
  wr<1>(a) /= s;
  return a;
}
template<Vec3 A>
inline A &operator/=(A &a, typename vec_traits<A>::scalar_type s)
{
  wr<0>(a) /= s;
  wr<1>(a) /= s;
  wr<2>(a) /= s;
  return a;
}

// ---- dot / cross / mag / normalized ----
template<Vec2 A, Vec2 B>
inline auto dot(A const &a, B const &b)
{
  return rd<0>(a) * rd<0>(b) + rd<1>(a) * rd<1>(b);
}
template<Vec3 A, Vec3 B>
inline auto dot(A const &a, B const &b)
{
  return rd<0>(a) * rd<0>(b) + rd<1>(a) * rd<1>(b) + rd<2>(a) * rd<2>(b);
}
template<Vec3 A, Vec3 B>
inline typename deduce_vec2<A, B, 3>::type cross(A const &a, B const &b)
{
  typename deduce_vec2<A, B, 3>::type r;
  wr<0>(r) = rd<1>(a) * rd<2>(b) - rd<2>(a) * rd<1>(b);
  wr<1>(r) = rd<2>(a) * rd<0>(b) - rd<0>(a) * rd<2>(b);
  wr<2>(r) = rd<0>(a) * rd<1>(b) - rd<1>(a) * rd<0>(b);
  return r;
}
template<Vec2 A>
inline auto mag(A const &a)
{
  return std::sqrt(rd<0>(a) * rd<0>(a) + rd<1>(a) * rd<1>(a));
}
template<Vec3 A>
inline auto mag(A const &a)
{
  return std::sqrt(rd<0>(a) * rd<0>(a) + rd<1>(a) * rd<1>(a) + rd<2>(a) * rd<2>(a));
}
template<Vec2 A>
inline A normalized(A const &a)
{
  const auto m2 = rd<0>(a) * rd<0>(a) + rd<1>(a) * rd<1>(a);
  if(m2 == 0)
    throw std::runtime_error("zero magnitude");
  const auto rm = 1 / std::sqrt(m2);
  A r;
  wr<0>(r) = rd<0>(a) * rm;
  wr<1>(r) = rd<1>(a) * rm;
  return r;
}
template<Vec3 A>
inline A normalized(A const &a)
{
  const auto m2 = rd<0>(a) * rd<0>(a) + rd<1>(a) * rd<1>(a) + rd<2>(a) * rd<2>(a);
  if(m2 == 0)
    throw std::runtime_error("zero magnitude");
  const auto rm = 1 / std::sqrt(m2);
  A r;
  wr<0>(r) = rd<0>(a) * rm;
  wr<1>(r) = rd<1>(a) * rm;
  wr<2>(r) = rd<2>(a) * rm;
  return r;
}

// ---- matrices ----
template<class T, int D>
inline mat<T, D, D> identity_mat()
{
  mat<T, D, D> m;
  for(int i = 0; i < D; i++)
    for(int j = 0; j < D; j++)
      m.a[i][j] = i == j ? T(1) : T(0);
  return m;
}
template<class T, int D>
inline mat<T, D, D> transposed(mat<T, D, D> const &m)
{
  mat<T, D, D> r;
  for(int i = 0; i < D; i++)
    for(int j = 0; j < D; j++)
      r.a[i][j] = m.a[j][i];
  return r;
}
template<class T>
inline T determinant(mat<T, 2, 2> const &m)
{
  return m.a[0][0] * m.a[1][1] - m.a[0][1] * m.a[1][0];
}
template<class T>
inline T determinant(mat<T, 3, 3> const &m)
{
  return m.a[0][0] * (m.a[1][1] * m.a[2][2] - m.a[1][2] * m.a[2][1])
       - m.a[0][1] * (m.a[1][0] * m.a[2][2] - m.a[1][2] * m.a[2][0])
       + m.a[0][2] * (m.a[1][0] * m.a[2][1] - m.a[1][1] * m.a[2][0]);
}
template<class T>
inline mat<T, 3, 3> inverse(mat<T, 3, 3> const &m)
{
  const T det = determinant(m);
  if(det == 0)
    throw std::runtime_error("zero determinant");
  const T f = 1 / det;
  mat<T, 3, 3> r;
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
template<class T>
inline mat<T, 2, 2> operator*(mat<T, 2, 2> const &a, mat<T, 2, 2> const &b)
{
  mat<T, 2, 2> r;
  for(int i = 0; i < 2; i++)
    for(int j = 0; j < 2; j++)
      r.a[i][j] = a.a[i][0] * b.a[0][j] + a.a[i][1] * b.a[1][j];
  return r;
}
template<class T>
inline mat<T, 3, 3> operator*(mat<T, 3, 3> const &a, mat<T, 3, 3> const &b)
{
  mat<T, 3, 3> r;
  for(int i = 0; i < 3; i++)
    for(int j = 0; j < 3; j++)
      r.a[i][j] = a.a[i][0] * b.a[0][j] + a.a[i][1] * b.a[1][j] + a.a[i][2] * b.a[2][j];
  return r;
}

// mat * vec: r_i = a_i0*b0 + a_i1*b1 (+ a_i2*b2), left to right (boost/qvm/gen/vec_mat_operations{2,3}.hpp)
template<class T, Vec2 B>
inline typename deduce_vec2<mat<T, 2, 2>, B, 2>::type operator*(mat<T, 2, 2> const &a, B const &b)
{
  typename deduce_vec2<mat<T, 2, 2>, B, 2>::type r;
  const auto b0 = rd<0>(b);
  const auto b1 = rd<1>(b);
  wr<0>(r) = a.a[0][0] * b0 + a.a[0][1] * b1;
  wr<1>(r) = a.a[1][0] * b0 + a.a[1][1] * b1;
  return r;
}
template<class T, Vec3 B>
inline typename deduce_vec2<mat<T, 3, 3>, B, 3>::type operator*(mat<T, 3, 3> const &a, B const &b)
{
  typename deduce_vec2<mat<T, 3, 3>, B, 3>::type r;
  const auto b0 = rd<0>(b);
  const auto b1 = rd<1>(b);
  const auto b2 = rd<2>(b);
  wr<0>(r) = a.a[0][0] * b0 + a.a[0][1] * b1 + a.a[0][2] * b2;
  wr<1>(r) = a.a[1][0] * b0 + a.a[1][1] * b1 + a.a[1][2] * b2;
  wr<2>(r) = a.a[2][0] * b0 + a.a[2][1] * b1 + a.a[2][2] * b2;
  return r;
}

// col<N>(m) = v
template<int N, class T, int D>
struct col_ref
{
  mat<T, D, D> &m;
  template<class V>
  col_ref &operator=(V const &v)
  {
    static_assert(vec_traits<V>::dim == D, "dimension mismatch");
    m.a[0][N] = rd<0>(v);
    m.a[1][N] = rd<1>(v);
    if constexpr(D == 3)
      m.a[2][N] = rd<2>(v);
    return *this;
  }
};
template<int N, class T, int D>
inline col_ref<N, T, D> col(mat<T, D, D> &m)
{
  return col_ref<N, T, D>{ m };
}

}} // namespace boost::qvm

// GCC >= 12 no longer finds dependent operators at instantiation time (PR c++/51577); the reference's
// transformation.h:68 relies on the old lookup. Make operator- visible in namespace stairs.
namespace stairs
{
using boost::qvm::operator-;
}
#endif
