// mini_eigen.hpp - a small stand-in for the parts of Eigen 3.3 that the reference's rotation-averaging library
// (ral/l1_irls.hpp, ral/l1_irls.cpp, ral/test.cpp) uses, so that those files can be compiled UNMODIFIED from
// /root/reference in an image that has no Eigen (see oracle/ref_shim/README.md and oracle/build_ref.py).
//
// TEST INFRASTRUCTURE ONLY (oracle/): the product (irotavg_b200/, include/) never includes this.
//
// Eager evaluation instead of expression templates: every operator returns a value.  The element-wise
// arithmetic is the same IEEE double arithmetic in the order the reference's source spells it; only reductions
// (norm, dot, sparse products) may differ from Eigen's vectorised summation order by rounding.  Storage is
// column-major like Eigen's default.  Quaternion product / normalisation follow Eigen/src/Geometry/Quaternion.h
// (generic, non-SIMD path).
#ifndef ORACLE_REF_SHIM_MINI_EIGEN_HPP_
#define ORACLE_REF_SHIM_MINI_EIGEN_HPP_

#include <algorithm>
#include <cassert>
#include <cctype>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

#define EIGEN_PI 3.141592653589793238462643383279502884197169399375105820974944592307816406L

namespace Eigen {

typedef std::ptrdiff_t Index;
enum { ColMajor = 0, RowMajor = 1 };
enum { StreamPrecision = -1, FullPrecision = -2 };

struct IOFormat {
  int precision;
  explicit IOFormat(int p = StreamPrecision) : precision(p) {}
};

class Matrix;
class Array;
template <bool W> class ViewT;
template <bool W> class AViewT;
typedef ViewT<true> View;
typedef ViewT<false> CView;
typedef AViewT<true> AView;
typedef AViewT<false> CAView;

struct MTag {};   // matrix family
struct ATag {};   // array (coefficient-wise) family
template <class T> struct is_m : std::is_base_of<MTag, typename std::decay<T>::type> {};
template <class T> struct is_a : std::is_base_of<ATag, typename std::decay<T>::type> {};
template <class T> struct is_dense : std::integral_constant<bool, is_m<T>::value || is_a<T>::value> {};
#define ME_IF(c) typename std::enable_if<(c), int>::type = 0

// rows x cols read access for anything dense
template <class D> struct DenseBase {
  const D& derived() const { return *static_cast<const D*>(this); }
  D& derived() { return *static_cast<D*>(this); }
  Index rows() const { return derived().rows_(); }
  Index cols() const { return derived().cols_(); }
  Index size() const { return rows() * cols(); }
  double coeff(Index i, Index j) const { return derived().at(i, j); }
  double lin(Index k) const { const Index r = rows(); return r == 1 ? derived().at(0, k) : derived().at(k % r, k / r); }
};

// copy src into a writable dst; vectors may be transposed (Eigen allows row <-> column vector assignment)
template <class Dst, class Src> void assign_dense(Dst& dst, const Src& src) {
  const Index r = dst.rows(), c = dst.cols();
  if (src.rows() == r && src.cols() == c) {
    for (Index j = 0; j < c; ++j)
      for (Index i = 0; i < r; ++i) dst.ref(i, j) = src.coeff(i, j);
  } else {
    assert((r == 1 || c == 1) && (src.rows() == 1 || src.cols() == 1) && src.size() == r * c);
    for (Index k = 0; k < r * c; ++k) dst.lref(k) = src.lin(k);
  }
}

struct Formatted {
  std::vector<double> v; Index r, c; IOFormat fmt;
};
inline std::ostream& operator<<(std::ostream& s, const Formatted& f) {
  // Eigen/src/Core/IO.h print_matrix: precision (FullPrecision = digits10 = 15 for double), columns aligned to
  // the widest entry, coefficient separator " ", row separator "\n"
  std::streamsize explicit_precision = 0;
  if (f.fmt.precision == FullPrecision) explicit_precision = std::numeric_limits<double>::digits10;
  else if (f.fmt.precision != StreamPrecision) explicit_precision = f.fmt.precision;
  std::streamsize old = 0;
  if (explicit_precision) old = s.precision(explicit_precision);
  Index width = 0;
  for (Index j = 0; j < f.c; ++j)
    for (Index i = 0; i < f.r; ++i) {
      std::stringstream ss; ss.copyfmt(s); ss << f.v[(size_t)(j * f.r + i)];
      width = std::max<Index>(width, (Index)ss.str().length());
    }
  for (Index i = 0; i < f.r; ++i) {
    if (width) s.width(width);
    s << f.v[(size_t)i];
    for (Index j = 1; j < f.c; ++j) { s << " "; if (width) s.width(width); s << f.v[(size_t)(j * f.r + i)]; }
    if (i < f.r - 1) s << "\n";
  }
  if (explicit_precision) s.precision(old);
  return s;
}

template <class D> struct Rowwise;
template <class D> struct CommaInit;
class SparseVecView;

// ---- array family -------------------------------------------------------------------------------------
template <class D> struct ArrayBase : DenseBase<D>, ATag {
  using DenseBase<D>::derived; using DenseBase<D>::rows; using DenseBase<D>::cols; using DenseBase<D>::coeff;
  template <class F> Array map(F f) const;
  Array abs() const; Array inverse() const; Array square() const; Array sqrt() const; Array sin() const;
  Array cos() const; Array tanh() const; Array exp() const; Array pow(double p) const;
  const D& array() const { return derived(); }
  D& array() { return derived(); }
  template <class O, ME_IF(is_dense<O>::value)> D& operator*=(const O& o) { return zip(o, [](double a, double b) { return a * b; }); }
  template <class O, ME_IF(is_dense<O>::value)> D& operator+=(const O& o) { return zip(o, [](double a, double b) { return a + b; }); }
  template <class O, ME_IF(is_dense<O>::value)> D& operator-=(const O& o) { return zip(o, [](double a, double b) { return a - b; }); }
  template <class O, ME_IF(is_dense<O>::value)> D& operator/=(const O& o) { return zip(o, [](double a, double b) { return a / b; }); }
  D& operator*=(double s) { return each([s](double a) { return a * s; }); }
  D& operator/=(double s) { return each([s](double a) { return a / s; }); }
  D& operator+=(double s) { return each([s](double a) { return a + s; }); }
  D& operator-=(double s) { return each([s](double a) { return a - s; }); }
 private:
  template <class O, class F> D& zip(const O& o, F f) {
    D& d = derived();
    if (o.rows() == rows() && o.cols() == cols()) {
      for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) d.ref(i, j) = f(d.at(i, j), o.coeff(i, j));
    } else {
      assert(o.size() == this->size());
      for (Index k = 0; k < this->size(); ++k) d.lref(k) = f(this->lin(k), o.lin(k));
    }
    return d;
  }
  template <class F> D& each(F f) {
    D& d = derived();
    for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) d.ref(i, j) = f(d.at(i, j));
    return d;
  }
};

// ---- matrix family ------------------------------------------------------------------------------------
template <class D> struct MatrixBase : DenseBase<D>, MTag {
  using DenseBase<D>::derived; using DenseBase<D>::rows; using DenseBase<D>::cols; using DenseBase<D>::coeff;
  MatrixBase() {}
  MatrixBase(const MatrixBase&) {}
  // assignment through the base (ral/l1_irls.cpp:552 assigns a Map to a MatrixBase<Derived>&)
  MatrixBase& operator=(const MatrixBase& o) { assign_dense(derived(), o.derived()); return *this; }
  template <class O, ME_IF(is_dense<O>::value)> D& operator=(const O& o) { derived().assign(o); return derived(); }
  double squaredNorm() const { double s = 0; for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) { const double v = coeff(i, j); s += v * v; } return s; }
  double norm() const { return std::sqrt(squaredNorm()); }
  double sum() const { double s = 0; for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) s += coeff(i, j); return s; }
  double mean() const { return sum() / (double)this->size(); }
  double maxCoeff() const { double m = coeff(0, 0); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) m = std::max(m, coeff(i, j)); return m; }
  template <class O> double dot(const O& o) const { double s = 0; for (Index k = 0; k < this->size(); ++k) s += this->lin(k) * o.lin(k); return s; }
  Rowwise<D> rowwise() const;
  template <class F> Matrix unaryExpr(F f) const;
  Formatted format(const IOFormat& fmt) const;
  SparseVecView sparseView() const;
};

template <bool W> class ViewT : public MatrixBase<ViewT<W>> {
 public:
  typedef typename std::conditional<W, double, const double>::type S;
  S* p; Index r, c, rs, cs;
  ViewT(S* p_, Index r_, Index c_, Index rs_, Index cs_) : p(p_), r(r_), c(c_), rs(rs_), cs(cs_) {}
  ViewT(const ViewT& o) : MatrixBase<ViewT<W>>(), p(o.p), r(o.r), c(o.c), rs(o.rs), cs(o.cs) {}
  Index rows_() const { return r; }
  Index cols_() const { return c; }
  double at(Index i, Index j) const { return p[i * rs + j * cs]; }
  S& ref(Index i, Index j) const { return p[i * rs + j * cs]; }
  S& lref(Index k) const { return r == 1 ? p[k * cs] : p[(k % r) * rs + (k / r) * cs]; }
  S& operator()(Index i, Index j) const { return ref(i, j); }
  S& operator()(Index k) const { return lref(k); }
  template <class O> void assign(const O& o) { assign_dense(*this, o); }
  ViewT& operator=(const ViewT& o) { assign_dense(*this, o); return *this; }          // copies coefficients
  template <class O, ME_IF(is_dense<O>::value)> ViewT& operator=(const O& o) { assign_dense(*this, o); return *this; }
  ViewT& operator/=(double s) { for (Index j = 0; j < c; ++j) for (Index i = 0; i < r; ++i) ref(i, j) /= s; return *this; }
  ViewT& operator*=(double s) { for (Index j = 0; j < c; ++j) for (Index i = 0; i < r; ++i) ref(i, j) *= s; return *this; }
  void setZero() { for (Index j = 0; j < c; ++j) for (Index i = 0; i < r; ++i) ref(i, j) = 0.0; }
  void setOnes() { for (Index j = 0; j < c; ++j) for (Index i = 0; i < r; ++i) ref(i, j) = 1.0; }
  AViewT<W> array() const;
  // sub-views (a vector's head/tail run along its only dimension)
  ViewT head(Index n) const { return r == 1 ? ViewT(p, 1, n, rs, cs) : ViewT(p, n, 1, rs, cs); }
  ViewT tail(Index n) const { return r == 1 ? ViewT(p + (c - n) * cs, 1, n, rs, cs) : ViewT(p + (r - n) * rs, n, 1, rs, cs); }
  ViewT row(Index i) const { return ViewT(p + i * rs, 1, c, rs, cs); }
  ViewT col(Index j) const { return ViewT(p + j * cs, r, 1, rs, cs); }
  ViewT leftCols(Index n) const { return ViewT(p, r, n, rs, cs); }
  CommaInit<ViewT> operator<<(double v);
};

template <bool W> class AViewT : public ArrayBase<AViewT<W>> {
 public:
  typedef typename std::conditional<W, double, const double>::type S;
  S* p; Index r, c, rs, cs;
  AViewT(S* p_, Index r_, Index c_, Index rs_, Index cs_) : p(p_), r(r_), c(c_), rs(rs_), cs(cs_) {}
  Index rows_() const { return r; }
  Index cols_() const { return c; }
  double at(Index i, Index j) const { return p[i * rs + j * cs]; }
  S& ref(Index i, Index j) const { return p[i * rs + j * cs]; }
  S& lref(Index k) const { return r == 1 ? p[k * cs] : p[(k % r) * rs + (k / r) * cs]; }
  using ArrayBase<AViewT<W>>::operator*=; using ArrayBase<AViewT<W>>::operator+=;
  using ArrayBase<AViewT<W>>::operator-=; using ArrayBase<AViewT<W>>::operator/=;
};
template <bool W> AViewT<W> ViewT<W>::array() const { return AViewT<W>(p, r, c, rs, cs); }

class Array : public ArrayBase<Array> {
 public:
  std::vector<double> d; Index r = 0, c = 0;
  Array() {}
  Array(Index r_, Index c_) : d((size_t)(r_ * c_)), r(r_), c(c_) {}
  template <class O, ME_IF(is_dense<O>::value)> Array(const O& o) : d((size_t)o.size()), r(o.rows()), c(o.cols()) {
    for (Index j = 0; j < c; ++j) for (Index i = 0; i < r; ++i) d[(size_t)(j * r + i)] = o.coeff(i, j);
  }
  Index rows_() const { return r; }
  Index cols_() const { return c; }
  double at(Index i, Index j) const { return d[(size_t)(j * r + i)]; }
  double& ref(Index i, Index j) { return d[(size_t)(j * r + i)]; }
  double& lref(Index k) { return d[(size_t)k]; }
};

class Matrix : public MatrixBase<Matrix> {
 public:
  std::vector<double> d; Index r = 0, c = 0;
  Matrix() {}
  explicit Matrix(Index n) : d((size_t)n), r(n), c(1) {}
  explicit Matrix(int n) : d((size_t)n), r(n), c(1) {}
  Matrix(Index r_, Index c_) : d((size_t)(r_ * c_)), r(r_), c(c_) {}
  Matrix(double x, double y, double z, double w) : d{x, y, z, w}, r(4), c(1) {}
  Matrix(const Matrix& o) : MatrixBase<Matrix>(), d(o.d), r(o.r), c(o.c) {}
  template <class O, ME_IF(is_dense<O>::value)> Matrix(const O& o) : d((size_t)o.size()), r(o.rows()), c(o.cols()) {
    for (Index j = 0; j < c; ++j) for (Index i = 0; i < r; ++i) d[(size_t)(j * r + i)] = o.coeff(i, j);
  }
  static Matrix Zero(Index r_, Index c_) { Matrix m(r_, c_); std::fill(m.d.begin(), m.d.end(), 0.0); return m; }
  Index rows_() const { return r; }
  Index cols_() const { return c; }
  double at(Index i, Index j) const { return d[(size_t)(j * r + i)]; }
  double& ref(Index i, Index j) { return d[(size_t)(j * r + i)]; }
  double& lref(Index k) { return d[(size_t)k]; }
  double& operator()(Index i, Index j) { return ref(i, j); }
  double operator()(Index i, Index j) const { return at(i, j); }
  double& operator()(Index k) { return d[(size_t)k]; }
  double operator()(Index k) const { return d[(size_t)k]; }
  template <class O> void assign(const O& o) {            // resizing assignment
    Matrix t(o);
    d.swap(t.d); r = t.r; c = t.c;
  }
  Matrix& operator=(const Matrix& o) { d = o.d; r = o.r; c = o.c; return *this; }
  template <class O, ME_IF(is_dense<O>::value)> Matrix& operator=(const O& o) { assign(o); return *this; }
  double* data() { return d.data(); }
  const double* data() const { return d.data(); }
  Index outerStride() const { return r; }
  void setZero() { std::fill(d.begin(), d.end(), 0.0); }
  void setOnes() { std::fill(d.begin(), d.end(), 1.0); }
  View v() { return View(d.data(), r, c, 1, r); }
  CView v() const { return CView(d.data(), r, c, 1, r); }
  AView array() { return AView(d.data(), r, c, 1, r); }
  CAView array() const { return CAView(d.data(), r, c, 1, r); }
  View head(Index n) { return v().head(n); }
  CView head(Index n) const { return v().head(n); }
  View tail(Index n) { return v().tail(n); }
  CView tail(Index n) const { return v().tail(n); }
  View row(Index i) { return v().row(i); }
  CView row(Index i) const { return v().row(i); }
  View col(Index j) { return v().col(j); }
  CView col(Index j) const { return v().col(j); }
  View leftCols(Index n) { return v().leftCols(n); }
  CView leftCols(Index n) const { return v().leftCols(n); }
  Matrix& operator/=(double s) { for (double& x : d) x /= s; return *this; }
  Matrix& operator*=(double s) { for (double& x : d) x *= s; return *this; }
  void transposeInPlace() {
    Matrix t(c, r);
    for (Index j = 0; j < c; ++j) for (Index i = 0; i < r; ++i) t.ref(j, i) = at(i, j);
    d.swap(t.d); std::swap(r, c);
  }
};

// Eigen::Map<Mat>: a column-major view over caller memory
template <class M> class Map;
template <> class Map<Matrix> : public MatrixBase<Map<Matrix>> {
 public:
  double* p; Index r, c;
  Map(double* p_, Index r_, Index c_) : p(p_), r(r_), c(c_) {}
  Map(const Map& o) : MatrixBase<Map<Matrix>>(), p(o.p), r(o.r), c(o.c) {}
  Index rows_() const { return r; }
  Index cols_() const { return c; }
  double at(Index i, Index j) const { return p[j * r + i]; }
  double& ref(Index i, Index j) { return p[j * r + i]; }
  double& lref(Index k) { return p[k]; }
  template <class O> void assign(const O& o) { assign_dense(*this, o); }
  Map& operator=(const Map& o) { assign_dense(*this, o); return *this; }
  template <class O, ME_IF(is_dense<O>::value)> Map& operator=(const O& o) { assign_dense(*this, o); return *this; }
  CView v() const { return CView(p, r, c, 1, r); }
};

template <class D> struct CommaInit {
  D dst; Index k;
  CommaInit(const D& d, double v) : dst(d), k(0) { dst.lref(k++) = v; }
  CommaInit& operator,(double v) { dst.lref(k++) = v; return *this; }
};
template <bool W> CommaInit<ViewT<W>> ViewT<W>::operator<<(double v) { return CommaInit<ViewT<W>>(*this, v); }

template <class D> struct Rowwise {
  const D& m;
  explicit Rowwise(const D& m_) : m(m_) {}
  Matrix squaredNorm() const {
    Matrix o(m.rows(), (Index)1);
    for (Index i = 0; i < m.rows(); ++i) { double s = 0; for (Index j = 0; j < m.cols(); ++j) { const double v = m.coeff(i, j); s += v * v; } o.d[(size_t)i] = s; }
    return o;
  }
  Matrix norm() const { Matrix o = squaredNorm(); for (double& x : o.d) x = std::sqrt(x); return o; }
};
template <class D> Rowwise<D> MatrixBase<D>::rowwise() const { return Rowwise<D>(derived()); }

template <class D> template <class F> Matrix MatrixBase<D>::unaryExpr(F f) const {
  Matrix o(rows(), cols());
  for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) o.ref(i, j) = f(coeff(i, j));
  return o;
}
template <class D> Formatted MatrixBase<D>::format(const IOFormat& fmt) const {
  Formatted f; f.r = rows(); f.c = cols(); f.fmt = fmt; f.v.resize((size_t)(f.r * f.c));
  for (Index j = 0; j < f.c; ++j) for (Index i = 0; i < f.r; ++i) f.v[(size_t)(j * f.r + i)] = coeff(i, j);
  return f;
}

template <class D> template <class F> Array ArrayBase<D>::map(F f) const {
  Array o(rows(), cols());
  for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) o.ref(i, j) = f(coeff(i, j));
  return o;
}
template <class D> Array ArrayBase<D>::abs() const { return map([](double a) { return std::abs(a); }); }
template <class D> Array ArrayBase<D>::inverse() const { return map([](double a) { return 1.0 / a; }); }
template <class D> Array ArrayBase<D>::square() const { return map([](double a) { return a * a; }); }
template <class D> Array ArrayBase<D>::sqrt() const { return map([](double a) { return std::sqrt(a); }); }
template <class D> Array ArrayBase<D>::sin() const { return map([](double a) { return std::sin(a); }); }
template <class D> Array ArrayBase<D>::cos() const { return map([](double a) { return std::cos(a); }); }
template <class D> Array ArrayBase<D>::tanh() const { return map([](double a) { return std::tanh(a); }); }
template <class D> Array ArrayBase<D>::exp() const { return map([](double a) { return std::exp(a); }); }
template <class D> Array ArrayBase<D>::pow(double p) const { return map([p](double a) { return std::pow(a, p); }); }

// ---- operators ------------------------------------------------------------------------------------------
template <class R, class A, class B, class F> R zip2(const A& a, const B& b, F f) {
  R o(a.rows(), a.cols());
  if (a.rows() == b.rows() && a.cols() == b.cols()) {
    for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) o.ref(i, j) = f(a.coeff(i, j), b.coeff(i, j));
  } else {
    assert(a.size() == b.size());
    for (Index k = 0; k < a.size(); ++k) o.lref(k) = f(a.lin(k), b.lin(k));
  }
  return o;
}
template <class R, class A, class F> R map1(const A& a, F f) {
  R o(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) o.ref(i, j) = f(a.coeff(i, j));
  return o;
}
// matrix family
template <class A, class B, ME_IF(is_m<A>::value && is_m<B>::value)> Matrix operator+(const A& a, const B& b) { return zip2<Matrix>(a, b, [](double x, double y) { return x + y; }); }
template <class A, class B, ME_IF(is_m<A>::value && is_m<B>::value)> Matrix operator-(const A& a, const B& b) { return zip2<Matrix>(a, b, [](double x, double y) { return x - y; }); }
template <class A, ME_IF(is_m<A>::value)> Matrix operator-(const A& a) { return map1<Matrix>(a, [](double x) { return -x; }); }
template <class A, ME_IF(is_m<A>::value)> Matrix operator*(const A& a, double s) { return map1<Matrix>(a, [s](double x) { return x * s; }); }
template <class A, ME_IF(is_m<A>::value)> Matrix operator*(double s, const A& a) { return map1<Matrix>(a, [s](double x) { return s * x; }); }
template <class A, ME_IF(is_m<A>::value)> Matrix operator/(const A& a, double s) { return map1<Matrix>(a, [s](double x) { return x / s; }); }
// array family
template <class A, class B, ME_IF(is_a<A>::value && is_a<B>::value)> Array operator+(const A& a, const B& b) { return zip2<Array>(a, b, [](double x, double y) { return x + y; }); }
template <class A, class B, ME_IF(is_a<A>::value && is_a<B>::value)> Array operator-(const A& a, const B& b) { return zip2<Array>(a, b, [](double x, double y) { return x - y; }); }
template <class A, class B, ME_IF(is_a<A>::value && is_a<B>::value)> Array operator*(const A& a, const B& b) { return zip2<Array>(a, b, [](double x, double y) { return x * y; }); }
template <class A, class B, ME_IF(is_a<A>::value && is_a<B>::value)> Array operator/(const A& a, const B& b) { return zip2<Array>(a, b, [](double x, double y) { return x / y; }); }
template <class A, ME_IF(is_a<A>::value)> Array operator-(const A& a) { return map1<Array>(a, [](double x) { return -x; }); }
template <class A, ME_IF(is_a<A>::value)> Array operator+(const A& a, double s) { return map1<Array>(a, [s](double x) { return x + s; }); }
template <class A, ME_IF(is_a<A>::value)> Array operator+(double s, const A& a) { return map1<Array>(a, [s](double x) { return s + x; }); }
template <class A, ME_IF(is_a<A>::value)> Array operator-(const A& a, double s) { return map1<Array>(a, [s](double x) { return x - s; }); }
template <class A, ME_IF(is_a<A>::value)> Array operator-(double s, const A& a) { return map1<Array>(a, [s](double x) { return s - x; }); }
template <class A, ME_IF(is_a<A>::value)> Array operator*(const A& a, double s) { return map1<Array>(a, [s](double x) { return x * s; }); }
template <class A, ME_IF(is_a<A>::value)> Array operator*(double s, const A& a) { return map1<Array>(a, [s](double x) { return s * x; }); }
template <class A, ME_IF(is_a<A>::value)> Array operator/(const A& a, double s) { return map1<Array>(a, [s](double x) { return x / s; }); }

typedef Matrix MatrixXd;
typedef Matrix VectorXd;
typedef Matrix Vector3d;
typedef Matrix Vector4d;
typedef Matrix Matrix3d;

// ---- Quaternion (Eigen/src/Geometry/Quaternion.h, generic path) ----------------------------------------
class Quaterniond {
 public:
  double x_, y_, z_, w_;
  Quaterniond() : x_(0), y_(0), z_(0), w_(1) {}
  Quaterniond(double w, double x, double y, double z) : x_(x), y_(y), z_(z), w_(w) {}
  double x() const { return x_; } double y() const { return y_; } double z() const { return z_; } double w() const { return w_; }
  Quaterniond operator*(const Quaterniond& b) const {
    const Quaterniond& a = *this;
    return Quaterniond(a.w_ * b.w_ - a.x_ * b.x_ - a.y_ * b.y_ - a.z_ * b.z_,
                       a.w_ * b.x_ + a.x_ * b.w_ + a.y_ * b.z_ - a.z_ * b.y_,
                       a.w_ * b.y_ + a.y_ * b.w_ + a.z_ * b.x_ - a.x_ * b.z_,
                       a.w_ * b.z_ + a.z_ * b.w_ + a.x_ * b.y_ - a.y_ * b.x_);
  }
  Quaterniond& operator*=(const Quaterniond& b) { *this = *this * b; return *this; }
  double squaredNorm() const { return x_ * x_ + y_ * y_ + z_ * z_ + w_ * w_; }
  Quaterniond normalized() const {           // MatrixBase::normalized(): divide by the norm when it is > 0
    const double z2 = squaredNorm();
    if (z2 > 0.0) { const double n = std::sqrt(z2); return Quaterniond(w_ / n, x_ / n, y_ / n, z_ / n); }
    return *this;
  }
  void normalize() { *this = normalized(); }
  Matrix toRotationMatrix() const {         // Eigen/src/Geometry/Quaternion.h QuaternionBase::toRotationMatrix
    Matrix res(3, 3);
    const double tx = 2 * x_, ty = 2 * y_, tz = 2 * z_;
    const double twx = tx * w_, twy = ty * w_, twz = tz * w_, txx = tx * x_, txy = ty * x_, txz = tz * x_;
    const double tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
    res(0, 0) = 1 - (tyy + tzz); res(0, 1) = txy - twz; res(0, 2) = txz + twy;
    res(1, 0) = txy + twz; res(1, 1) = 1 - (txx + tzz); res(1, 2) = tyz - twx;
    res(2, 0) = txz - twy; res(2, 1) = tyz + twx; res(2, 2) = 1 - (txx + tyy);
    return res;
  }
};

// ---- sparse ---------------------------------------------------------------------------------------------
template <class S> class Triplet {
 public:
  Index r, c; S v;
  Triplet() : r(0), c(0), v(0) {}
  Triplet(Index r_, Index c_, const S& v_ = S(0)) : r(r_), c(c_), v(v_) {}
  Index row() const { return r; } Index col() const { return c; } const S& value() const { return v; }
};

template <class S, int Opt, class I> class SparseMatrix;
template <class M> class Ref;
template <class I> struct SparseTransposed;

// column-major compressed storage; while being filled through coeffRef, per-column sorted vectors
template <class I> class SparseMatrix<double, ColMajor, I> {
 public:
  typedef I StorageIndex;
  Index nr = 0, nc = 0;
  std::vector<I> outer;          // nc + 1
  std::vector<I> inner;
  std::vector<double> val;
  std::vector<std::vector<std::pair<I, double>>> build;   // uncompressed insertion mode
  bool compressed = true;
  SparseMatrix() : outer(1, 0) {}
  SparseMatrix(Index r, Index c) : nr(r), nc(c), outer((size_t)c + 1, 0) {}
  Index rows() const { return nr; }
  Index cols() const { return nc; }
  Index nonZeros() const { return (Index)val.size(); }
  bool isCompressed() const { return true; }
  I* outerIndexPtr() { finish(); return outer.data(); }
  I* innerIndexPtr() { finish(); return inner.data(); }
  double* valuePtr() { finish(); return val.data(); }
  I* innerNonZeroPtr() { return nullptr; }
  const I* outerIndexPtr() const { return outer.data(); }
  const I* innerIndexPtr() const { return inner.data(); }
  const double* valuePtr() const { return val.data(); }
  double& coeffRef(Index r, Index c) {
    if (compressed) {                                       // switch to insertion mode
      build.assign((size_t)nc, std::vector<std::pair<I, double>>());
      for (Index j = 0; j < nc; ++j)
        for (I k = outer[(size_t)j]; k < outer[(size_t)j + 1]; ++k) build[(size_t)j].push_back(std::make_pair(inner[(size_t)k], val[(size_t)k]));
      compressed = false;
    }
    std::vector<std::pair<I, double>>& col = build[(size_t)c];
    auto it = std::lower_bound(col.begin(), col.end(), (I)r, [](const std::pair<I, double>& a, I b) { return a.first < b; });
    if (it == col.end() || it->first != (I)r) it = col.insert(it, std::make_pair((I)r, 0.0));
    return it->second;
  }
  void makeCompressed() { finish(); }
  void finish() {
    if (compressed) return;
    inner.clear(); val.clear();
    for (Index j = 0; j < nc; ++j) {
      outer[(size_t)j] = (I)inner.size();
      for (auto& e : build[(size_t)j]) { inner.push_back(e.first); val.push_back(e.second); }
    }
    outer[(size_t)nc] = (I)inner.size();
    build.clear();
    compressed = true;
  }
  // duplicates are summed, like Eigen's setFromTriplets
  template <class It> void setFromTriplets(It b, It e) {
    std::vector<std::map<I, double>> cols((size_t)nc);
    for (It t = b; t != e; ++t) cols[(size_t)t->col()][(I)t->row()] += t->value();
    inner.clear(); val.clear(); outer.assign((size_t)nc + 1, 0);
    for (Index j = 0; j < nc; ++j) {
      outer[(size_t)j] = (I)inner.size();
      for (auto& kv : cols[(size_t)j]) { inner.push_back(kv.first); val.push_back(kv.second); }
    }
    outer[(size_t)nc] = (I)inner.size();
    compressed = true;
  }
  SparseTransposed<I> transpose() const;
};
template <class I> struct SparseTransposed { const SparseMatrix<double, ColMajor, I>& m; };
template <class I> SparseTransposed<I> SparseMatrix<double, ColMajor, I>::transpose() const { return SparseTransposed<I>{*this}; }

template <class I> class Ref<SparseMatrix<double, ColMajor, I>> {
 public:
  SparseMatrix<double, ColMajor, I>& m;
  Ref(SparseMatrix<double, ColMajor, I>& m_) : m(m_) { m.finish(); }
  Index rows() const { return m.rows(); } Index cols() const { return m.cols(); } Index nonZeros() const { return m.nonZeros(); }
  bool isCompressed() const { return true; }
  I* outerIndexPtr() { return m.outerIndexPtr(); } I* innerIndexPtr() { return m.innerIndexPtr(); }
  double* valuePtr() { return m.valuePtr(); } I* innerNonZeroPtr() { return nullptr; }
};

// sparse * dense -> dense (column by column, entries in storage order)
template <class I, class D, ME_IF(is_m<D>::value)>
Matrix operator*(const SparseMatrix<double, ColMajor, I>& A, const D& x) {
  assert(A.compressed && x.rows() == A.cols());
  Matrix y = Matrix::Zero(A.rows(), x.cols());
  for (Index c = 0; c < x.cols(); ++c)
    for (Index j = 0; j < A.cols(); ++j) {
      const double xj = x.coeff(j, c);
      for (I k = A.outer[(size_t)j]; k < A.outer[(size_t)j + 1]; ++k) y.ref((Index)A.inner[(size_t)k], c) += A.val[(size_t)k] * xj;
    }
  return y;
}
// sparse^T * dense -> dense
template <class I, class D, ME_IF(is_m<D>::value)>
Matrix operator*(const SparseTransposed<I>& At, const D& x) {
  const SparseMatrix<double, ColMajor, I>& A = At.m;
  assert(A.compressed && x.rows() == A.rows());
  Matrix y = Matrix::Zero(A.cols(), x.cols());
  for (Index c = 0; c < x.cols(); ++c)
    for (Index j = 0; j < A.cols(); ++j) {
      double s = 0.0;
      for (I k = A.outer[(size_t)j]; k < A.outer[(size_t)j + 1]; ++k) s += A.val[(size_t)k] * x.coeff((Index)A.inner[(size_t)k], c);
      y.ref(j, c) = s;
    }
  return y;
}
// sparse * sparse -> sparse with sorted inner indices (what Eigen's conservative product delivers into a
// column-major destination); structural entries are kept even when their value is 0
template <class I>
SparseMatrix<double, ColMajor, I> operator*(const SparseMatrix<double, ColMajor, I>& A, const SparseMatrix<double, ColMajor, I>& B) {
  assert(A.compressed && B.compressed && A.cols() == B.rows());
  SparseMatrix<double, ColMajor, I> C(A.rows(), B.cols());
  std::map<I, double> acc;
  for (Index j = 0; j < B.cols(); ++j) {
    acc.clear();
    for (I kb = B.outer[(size_t)j]; kb < B.outer[(size_t)j + 1]; ++kb) {
      const Index p = (Index)B.inner[(size_t)kb];
      const double bv = B.val[(size_t)kb];
      for (I ka = A.outer[(size_t)p]; ka < A.outer[(size_t)p + 1]; ++ka) acc[A.inner[(size_t)ka]] += A.val[(size_t)ka] * bv;
    }
    C.outer[(size_t)j] = (I)C.inner.size();
    for (auto& kv : acc) { C.inner.push_back(kv.first); C.val.push_back(kv.second); }
  }
  C.outer[(size_t)B.cols()] = (I)C.inner.size();
  return C;
}

// Vec::sparseView(): the non-zero coefficients of a dense vector as an (n x 1) sparse matrix
class SparseVecView {
 public:
  std::vector<double> v;
  template <class I> operator SparseMatrix<double, ColMajor, I>() const {
    SparseMatrix<double, ColMajor, I> S((Index)v.size(), 1);
    for (size_t k = 0; k < v.size(); ++k) if (v[k] != 0.0) { S.inner.push_back((I)k); S.val.push_back(v[k]); }
    S.outer[1] = (I)S.inner.size();
    return S;
  }
};
template <class D> SparseVecView MatrixBase<D>::sparseView() const {
  SparseVecView s; s.v.resize((size_t)this->size());
  for (Index k = 0; k < this->size(); ++k) s.v[(size_t)k] = this->lin(k);
  return s;
}

}  // namespace Eigen
#endif
