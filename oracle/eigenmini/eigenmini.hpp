// eigenmini -- TEST INFRASTRUCTURE (oracle/), never linked into the product library.
//
// A functional stand-in for the part of Eigen 3 that the reference's vendored g2o (Thirdparty/g2o/g2o/{core,types,solvers}),
// src/Optimizer.cc and src/Converter.cc use, so that those files compile UNMODIFIED, from where they lie under /root/reference,
// in an image that has no Eigen (see oracle/Makefile: `ref`).  Same role as oracle/cvmini for OpenCV.
//
// Everything is evaluated eagerly (no expression templates): an operator returns a plain Matrix.  The arithmetic that decides
// results is written the way Eigen's published algorithms do it:
//   * fixed-size products: coefficient (i,j) = sum over k in ascending order, one rounding per operation (no FMA; the oracle
//     libraries are built with -ffp-contract=off);
//   * 2x2 / 3x3 / 4x4 inverse and determinant: cofactor formulas (Eigen's compute_inverse_size{2,3,4}_helper); larger: partial-
//     pivot LU;
//   * Quaternion(Matrix3): Shoemake's branch on the trace; toRotationMatrix, q*v (v + w*uv + q.vec x uv, uv = 2 q.vec x v), q*q;
//   * dense LDLT: diagonal pivoting on the largest |D_ii| (Eigen's LDLT), LLT: unpivoted Cholesky;
//   * SimplicialLDLT: up-looking sparse LDL^T on the permuted upper triangle (natural ordering: Eigen's AMD only changes the
//     elimination order, i.e. the result up to rounding).
// Column-major only (the reference never asks for RowMajor).
#ifndef ORBX_EIGENMINI_HPP
#define ORBX_EIGENMINI_HPP

#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <type_traits>
#include <vector>

#define EIGEN_WORLD_VERSION 3
#define EIGEN_MAJOR_VERSION 2
#define EIGEN_MINOR_VERSION 10
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW_IF(x)
#define EIGEN_DEFINE_STL_VECTOR_SPECIALIZATION(...)
#define EIGEN_STRONG_INLINE inline
#define EIGEN_ALIGN16
#define EIGEN_VERSION_AT_LEAST(x, y, z) (EIGEN_WORLD_VERSION > x || (EIGEN_WORLD_VERSION >= x && (EIGEN_MAJOR_VERSION > y || (EIGEN_MAJOR_VERSION >= y && EIGEN_MINOR_VERSION >= z))))

namespace Eigen {

typedef std::ptrdiff_t DenseIndex;
typedef DenseIndex Index;
const int Dynamic = -1;
const int Infinity = -1;
enum { ColMajor = 0, RowMajor = 1, AutoAlign = 0, DontAlign = 2 };
enum { Lower = 1, Upper = 2, UnitDiag = 4, ZeroDiag = 8, UnitLower = 5, UnitUpper = 6, StrictlyLower = 9, StrictlyUpper = 10, SelfAdjoint = 16 };
enum { Unaligned = 0, Aligned = 1 };
enum { AlignedBit = 0x40, LvalueBit = 0x20, DirectAccessBit = 0x40 << 1 };
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };
enum TransformTraits { Isometry = 0x1, Affine = 0x2, AffineCompact = 0x10 | Affine, Projective = 0x20 };
enum { ComputeEigenvectors = 0x80, EigenvaluesOnly = 0x40 };
inline void initParallel() {}
template <class T> using aligned_allocator = std::allocator<T>;

template <class D> struct traits;
template <class S, int R, int C, int Opt = 0, int MR = R, int MC = C> class Matrix;
template <class S, int R, int C> class Block;
template <class M, int MapOpt = Unaligned, class Stride = void> class Map;
template <class S> class Quaternion;
template <class S> class AngleAxis;
template <class D> class ArrayWrapper;
template <class D> class NoAlias;
template <class D, int UpLo> class DenseSelfAdjointView;
template <class M> class LLT;
template <class M> class LDLT;
template <class M> class PartialPivLU;

namespace internal {
template <int A, int B> struct pick { enum { value = (A != Dynamic) ? A : B }; };
template <class T> struct is_arith { enum { value = std::is_arithmetic<T>::value }; };
}  // namespace internal

// ------------------------------------------------------------------------------------------------------------------------
// MatrixBase: everything dense (Matrix, Map, Block) is a strided column-major view with data(), rows(), cols(), outerStride()
// ------------------------------------------------------------------------------------------------------------------------
template <class Derived> class MatrixBase {
public:
  typedef typename traits<Derived>::Scalar Scalar;
  typedef Scalar RealScalar;
  typedef typename traits<Derived>::Ref Ref;
  typedef Eigen::Index Index;
  enum {
    RowsAtCompileTime = traits<Derived>::Rows,
    ColsAtCompileTime = traits<Derived>::Cols,
    SizeAtCompileTime = (RowsAtCompileTime == Dynamic || ColsAtCompileTime == Dynamic) ? Dynamic : RowsAtCompileTime * ColsAtCompileTime,
    IsVectorAtCompileTime = (RowsAtCompileTime == 1 || ColsAtCompileTime == 1),
    Flags = AlignedBit | LvalueBit
  };
  typedef Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime> PlainObject;
  typedef Matrix<Scalar, ColsAtCompileTime, RowsAtCompileTime> TransposeReturnType;
  typedef Block<Scalar, RowsAtCompileTime, 1> ColXpr;
  typedef Block<const Scalar, RowsAtCompileTime, 1> ConstColXpr;

  Derived& derived() { return *static_cast<Derived*>(this); }
  const Derived& derived() const { return *static_cast<const Derived*>(this); }
  Index rows() const { return derived().rows(); }
  Index cols() const { return derived().cols(); }
  Index size() const { return rows() * cols(); }
  Index outerStride() const { return derived().outerStride(); }
  Index innerStride() const { return 1; }

  const Scalar& coeff(Index i, Index j) const { return derived().data()[i + j * derived().outerStride()]; }
  const Scalar& coeff(Index i) const { return cols() == 1 ? coeff(i, 0) : coeff(0, i); }
  Ref coeffRef(Index i, Index j) { return derived().data()[i + j * derived().outerStride()]; }
  Ref coeffRef(Index i) { return cols() == 1 ? coeffRef(i, 0) : coeffRef(0, i); }
  const Scalar& operator()(Index i, Index j) const { assert(i >= 0 && i < rows() && j >= 0 && j < cols()); return coeff(i, j); }
  Ref operator()(Index i, Index j) { assert(i >= 0 && i < rows() && j >= 0 && j < cols()); return coeffRef(i, j); }
  const Scalar& operator()(Index i) const { assert(i >= 0 && i < size()); return coeff(i); }
  Ref operator()(Index i) { assert(i >= 0 && i < size()); return coeffRef(i); }
  const Scalar& operator[](Index i) const { return coeff(i); }
  Ref operator[](Index i) { return coeffRef(i); }
  const Scalar& x() const { return coeff(0); }
  const Scalar& y() const { return coeff(1); }
  const Scalar& z() const { return coeff(2); }
  const Scalar& w() const { return coeff(3); }
  Ref x() { return coeffRef(0); }
  Ref y() { return coeffRef(1); }
  Ref z() { return coeffRef(2); }
  Ref w() { return coeffRef(3); }

  PlainObject eval() const { return PlainObject(*this); }

  // ---- assignment family (lvalues) ----
  template <class O> Derived& assign(const MatrixBase<O>& o) {
    derived().resizeLike(o.rows(), o.cols());
    if ((const void*)o.derived().data() == (const void*)derived().data() && o.outerStride() == outerStride()) return derived();
    for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) = o.coeff(i, j);
    return derived();
  }
  template <class O> Derived& operator+=(const MatrixBase<O>& o) {
    assert(rows() == o.rows() && cols() == o.cols());
    for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) += o.coeff(i, j);
    return derived();
  }
  template <class O> Derived& operator-=(const MatrixBase<O>& o) {
    assert(rows() == o.rows() && cols() == o.cols());
    for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) -= o.coeff(i, j);
    return derived();
  }
  Derived& operator*=(const Scalar& s) { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) *= s; return derived(); }
  Derived& operator/=(const Scalar& s) { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) /= s; return derived(); }
  template <class O> Derived& operator*=(const MatrixBase<O>& o) { return assign((*this) * o); }
  Derived& setConstant(const Scalar& s) { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) = s; return derived(); }
  Derived& fill(const Scalar& s) { return setConstant(s); }
  Derived& setZero() { return setConstant(Scalar(0)); }
  Derived& setOnes() { return setConstant(Scalar(1)); }
  Derived& setIdentity() { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) = (i == j) ? Scalar(1) : Scalar(0); return derived(); }
  Derived& setZero(Index n) { derived().resize(n); return setZero(); }
  Derived& setZero(Index r, Index c) { derived().resize(r, c); return setZero(); }
  Derived& setIdentity(Index r, Index c) { derived().resize(r, c); return setIdentity(); }
  NoAlias<Derived> noalias() { return NoAlias<Derived>(derived()); }
  template <class O> void swap(MatrixBase<O>& o) { PlainObject t(*this); assign(o); o.assign(t); }

  // ---- blocks ----
  typedef Block<Scalar, Dynamic, Dynamic> BlockXpr;
  typedef Block<const Scalar, Dynamic, Dynamic> ConstBlockXpr;
  Block<Scalar, Dynamic, Dynamic> block(Index i, Index j, Index r, Index c) { return Block<Scalar, Dynamic, Dynamic>(&coeffRef(i, j), r, c, outerStride()); }
  Block<const Scalar, Dynamic, Dynamic> block(Index i, Index j, Index r, Index c) const { return Block<const Scalar, Dynamic, Dynamic>(&coeff(i, j), r, c, outerStride()); }
  template <int R, int C> Block<Scalar, R, C> block(Index i, Index j) { return Block<Scalar, R, C>(&coeffRef(i, j), R, C, outerStride()); }
  template <int R, int C> Block<const Scalar, R, C> block(Index i, Index j) const { return Block<const Scalar, R, C>(&coeff(i, j), R, C, outerStride()); }
  template <int R, int C> Block<Scalar, R, C> block(Index i, Index j, Index r, Index c) { return Block<Scalar, R, C>(&coeffRef(i, j), r, c, outerStride()); }
  template <int R, int C> Block<const Scalar, R, C> block(Index i, Index j, Index r, Index c) const { return Block<const Scalar, R, C>(&coeff(i, j), r, c, outerStride()); }
  template <int R, int C> Block<Scalar, R, C> topLeftCorner() { return block<R, C>(0, 0); }
  template <int R, int C> Block<const Scalar, R, C> topLeftCorner() const { return block<R, C>(0, 0); }
  template <int R, int C> Block<Scalar, R, C> topRightCorner() { return block<R, C>(0, cols() - C); }
  template <int R, int C> Block<const Scalar, R, C> topRightCorner() const { return block<R, C>(0, cols() - C); }
  template <int R, int C> Block<Scalar, R, C> bottomLeftCorner() { return block<R, C>(rows() - R, 0); }
  template <int R, int C> Block<const Scalar, R, C> bottomLeftCorner() const { return block<R, C>(rows() - R, 0); }
  template <int R, int C> Block<Scalar, R, C> bottomRightCorner() { return block<R, C>(rows() - R, cols() - C); }
  template <int R, int C> Block<const Scalar, R, C> bottomRightCorner() const { return block<R, C>(rows() - R, cols() - C); }
  BlockXpr topLeftCorner(Index r, Index c) { return block(0, 0, r, c); }
  ConstBlockXpr topLeftCorner(Index r, Index c) const { return block(0, 0, r, c); }
  BlockXpr topRightCorner(Index r, Index c) { return block(0, cols() - c, r, c); }
  ConstBlockXpr topRightCorner(Index r, Index c) const { return block(0, cols() - c, r, c); }
  BlockXpr bottomLeftCorner(Index r, Index c) { return block(rows() - r, 0, r, c); }
  ConstBlockXpr bottomLeftCorner(Index r, Index c) const { return block(rows() - r, 0, r, c); }
  BlockXpr bottomRightCorner(Index r, Index c) { return block(rows() - r, cols() - c, r, c); }
  ConstBlockXpr bottomRightCorner(Index r, Index c) const { return block(rows() - r, cols() - c, r, c); }
  Block<Scalar, RowsAtCompileTime, 1> col(Index j) { return Block<Scalar, RowsAtCompileTime, 1>(&coeffRef(0, j), rows(), 1, outerStride()); }
  Block<const Scalar, RowsAtCompileTime, 1> col(Index j) const { return Block<const Scalar, RowsAtCompileTime, 1>(&coeff(0, j), rows(), 1, outerStride()); }
  Block<Scalar, 1, ColsAtCompileTime> row(Index i) { return Block<Scalar, 1, ColsAtCompileTime>(&coeffRef(i, 0), 1, cols(), outerStride()); }
  Block<const Scalar, 1, ColsAtCompileTime> row(Index i) const { return Block<const Scalar, 1, ColsAtCompileTime>(&coeff(i, 0), 1, cols(), outerStride()); }
  BlockXpr topRows(Index n) { return block(0, 0, n, cols()); }
  ConstBlockXpr topRows(Index n) const { return block(0, 0, n, cols()); }
  BlockXpr bottomRows(Index n) { return block(rows() - n, 0, n, cols()); }
  ConstBlockXpr bottomRows(Index n) const { return block(rows() - n, 0, n, cols()); }
  BlockXpr leftCols(Index n) { return block(0, 0, rows(), n); }
  ConstBlockXpr leftCols(Index n) const { return block(0, 0, rows(), n); }
  BlockXpr rightCols(Index n) { return block(0, cols() - n, rows(), n); }
  ConstBlockXpr rightCols(Index n) const { return block(0, cols() - n, rows(), n); }
  template <int N> Block<Scalar, N, ColsAtCompileTime> topRows() { return Block<Scalar, N, ColsAtCompileTime>(&coeffRef(0, 0), N, cols(), outerStride()); }
  template <int N> Block<const Scalar, N, ColsAtCompileTime> topRows() const { return Block<const Scalar, N, ColsAtCompileTime>(&coeff(0, 0), N, cols(), outerStride()); }
  template <int N> Block<Scalar, RowsAtCompileTime, N> leftCols() { return Block<Scalar, RowsAtCompileTime, N>(&coeffRef(0, 0), rows(), N, outerStride()); }
  template <int N> Block<const Scalar, RowsAtCompileTime, N> leftCols() const { return Block<const Scalar, RowsAtCompileTime, N>(&coeff(0, 0), rows(), N, outerStride()); }
  // vector segments: a column vector gives (n x 1), a row vector (1 x n)
  enum { SegR = (ColsAtCompileTime == 1) ? Dynamic : 1, SegC = (ColsAtCompileTime == 1) ? 1 : Dynamic };
  template <int N> struct Seg { enum { R = (ColsAtCompileTime == 1) ? N : 1, C = (ColsAtCompileTime == 1) ? 1 : N }; };
  Block<Scalar, SegR, SegC> segment(Index s, Index n) {
    return cols() == 1 ? Block<Scalar, SegR, SegC>(&coeffRef(s, 0), n, 1, outerStride()) : Block<Scalar, SegR, SegC>(&coeffRef(0, s), 1, n, outerStride());
  }
  Block<const Scalar, SegR, SegC> segment(Index s, Index n) const {
    return cols() == 1 ? Block<const Scalar, SegR, SegC>(&coeff(s, 0), n, 1, outerStride()) : Block<const Scalar, SegR, SegC>(&coeff(0, s), 1, n, outerStride());
  }
  template <int N> Block<Scalar, Seg<N>::R, Seg<N>::C> segment(Index s) {
    return cols() == 1 ? Block<Scalar, Seg<N>::R, Seg<N>::C>(&coeffRef(s, 0), N, 1, outerStride()) : Block<Scalar, Seg<N>::R, Seg<N>::C>(&coeffRef(0, s), 1, N, outerStride());
  }
  template <int N> Block<const Scalar, Seg<N>::R, Seg<N>::C> segment(Index s) const {
    return cols() == 1 ? Block<const Scalar, Seg<N>::R, Seg<N>::C>(&coeff(s, 0), N, 1, outerStride()) : Block<const Scalar, Seg<N>::R, Seg<N>::C>(&coeff(0, s), 1, N, outerStride());
  }
  template <int N> Block<Scalar, Seg<N>::R, Seg<N>::C> segment(Index s, Index) { return segment<N>(s); }
  template <int N> Block<const Scalar, Seg<N>::R, Seg<N>::C> segment(Index s, Index) const { return segment<N>(s); }
  Block<Scalar, SegR, SegC> head(Index n) { return segment(0, n); }
  Block<const Scalar, SegR, SegC> head(Index n) const { return segment(0, n); }
  Block<Scalar, SegR, SegC> tail(Index n) { return segment(size() - n, n); }
  Block<const Scalar, SegR, SegC> tail(Index n) const { return segment(size() - n, n); }
  template <int N> Block<Scalar, Seg<N>::R, Seg<N>::C> head() { return segment<N>(0); }
  template <int N> Block<const Scalar, Seg<N>::R, Seg<N>::C> head() const { return segment<N>(0); }
  template <int N> Block<Scalar, Seg<N>::R, Seg<N>::C> tail() { return segment<N>(size() - N); }
  template <int N> Block<const Scalar, Seg<N>::R, Seg<N>::C> tail() const { return segment<N>(size() - N); }
  // the diagonal as a strided view is not column-major; it is returned as a small proxy that supports what g2o writes
  class DiagonalProxy;
  class ConstDiagonal;
  DiagonalProxy diagonal() { return DiagonalProxy(derived()); }
  Matrix<Scalar, internal::pick<RowsAtCompileTime, ColsAtCompileTime>::value, 1> diagonal() const {
    Matrix<Scalar, internal::pick<RowsAtCompileTime, ColsAtCompileTime>::value, 1> d(std::min(rows(), cols()));
    for (Index i = 0; i < d.size(); ++i) d[i] = coeff(i, i);
    return d;
  }

  // ---- value-returning operations ----
  TransposeReturnType transpose() const {
    TransposeReturnType t(cols(), rows());
    for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) t.coeffRef(j, i) = coeff(i, j);
    return t;
  }
  TransposeReturnType adjoint() const { return transpose(); }
  void transposeInPlace() { TransposeReturnType t = transpose(); derived().resizeLike(t.rows(), t.cols()); assign(t); }
  PlainObject operator-() const { PlainObject r(rows(), cols()); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = -coeff(i, j); return r; }
  PlainObject cwiseAbs() const { PlainObject r(rows(), cols()); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = std::abs(coeff(i, j)); return r; }
  PlainObject cwiseAbs2() const { PlainObject r(rows(), cols()); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = coeff(i, j) * coeff(i, j); return r; }
  PlainObject cwiseSqrt() const { PlainObject r(rows(), cols()); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = std::sqrt(coeff(i, j)); return r; }
  PlainObject cwiseInverse() const { PlainObject r(rows(), cols()); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = Scalar(1) / coeff(i, j); return r; }
  template <class O> PlainObject cwiseProduct(const MatrixBase<O>& o) const { PlainObject r(rows(), cols()); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = coeff(i, j) * o.coeff(i, j); return r; }
  template <class O> PlainObject cwiseQuotient(const MatrixBase<O>& o) const { PlainObject r(rows(), cols()); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = coeff(i, j) / o.coeff(i, j); return r; }
  template <class O> PlainObject cwiseMax(const MatrixBase<O>& o) const { PlainObject r(rows(), cols()); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = std::max(coeff(i, j), o.coeff(i, j)); return r; }
  template <class O> PlainObject cwiseMin(const MatrixBase<O>& o) const { PlainObject r(rows(), cols()); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = std::min(coeff(i, j), o.coeff(i, j)); return r; }
  template <class T> Matrix<T, RowsAtCompileTime, ColsAtCompileTime> cast() const {
    Matrix<T, RowsAtCompileTime, ColsAtCompileTime> r(rows(), cols());
    for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = static_cast<T>(coeff(i, j));
    return r;
  }
  Scalar sum() const { Scalar s = Scalar(0); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) s += coeff(i, j); return s; }
  Scalar prod() const { Scalar s = Scalar(1); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) s *= coeff(i, j); return s; }
  Scalar mean() const { return sum() / Scalar(size()); }
  Scalar trace() const { Scalar s = Scalar(0); for (Index i = 0; i < std::min(rows(), cols()); ++i) s += coeff(i, i); return s; }
  Scalar squaredNorm() const { Scalar s = Scalar(0); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) s += coeff(i, j) * coeff(i, j); return s; }
  Scalar norm() const { return std::sqrt(squaredNorm()); }
  template <int P> Scalar lpNorm() const {
    if (P == Infinity) return cwiseAbs().maxCoeff();
    if (P == 1) return cwiseAbs().sum();
    return norm();
  }
  Scalar maxCoeff() const { Scalar m = coeff(0, 0); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (coeff(i, j) > m) m = coeff(i, j); return m; }
  Scalar minCoeff() const { Scalar m = coeff(0, 0); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (coeff(i, j) < m) m = coeff(i, j); return m; }
  template <class I> Scalar maxCoeff(I* idx) const { Scalar m = coeff(0); *idx = 0; for (Index i = 1; i < size(); ++i) if (coeff(i) > m) { m = coeff(i); *idx = I(i); } return m; }
  template <class I> Scalar minCoeff(I* idx) const { Scalar m = coeff(0); *idx = 0; for (Index i = 1; i < size(); ++i) if (coeff(i) < m) { m = coeff(i); *idx = I(i); } return m; }
  void normalize() { (*this) /= norm(); }
  PlainObject normalized() const { PlainObject r(*this); r /= norm(); return r; }
  template <class O> Scalar dot(const MatrixBase<O>& o) const { assert(size() == o.size()); Scalar s = Scalar(0); for (Index i = 0; i < size(); ++i) s += coeff(i) * o.coeff(i); return s; }
  template <class O> Matrix<Scalar, 3, 1> cross(const MatrixBase<O>& o) const {
    return Matrix<Scalar, 3, 1>(coeff(1) * o.coeff(2) - coeff(2) * o.coeff(1), coeff(2) * o.coeff(0) - coeff(0) * o.coeff(2), coeff(0) * o.coeff(1) - coeff(1) * o.coeff(0));
  }
  bool allFinite() const { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (!std::isfinite(coeff(i, j))) return false; return true; }
  bool hasNaN() const { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (std::isnan(coeff(i, j))) return true; return false; }
  template <class O> bool isApprox(const MatrixBase<O>& o, Scalar prec = Scalar(1e-12)) const { return ((*this) - o).squaredNorm() <= prec * prec * std::min(squaredNorm(), o.squaredNorm()); }
  template <class O> bool operator==(const MatrixBase<O>& o) const {
    if (rows() != o.rows() || cols() != o.cols()) return false;
    for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (!(coeff(i, j) == o.coeff(i, j))) return false;
    return true;
  }
  template <class O> bool operator!=(const MatrixBase<O>& o) const { return !(*this == o); }
  Scalar determinant() const;
  PlainObject inverse() const;
  ArrayWrapper<Derived> array() { return ArrayWrapper<Derived>(derived()); }
  PlainObject array() const { return PlainObject(*this); }
  PlainObject matrix() const { return PlainObject(*this); }
  template <int UpLo> DenseSelfAdjointView<Derived, UpLo> selfadjointView() { return DenseSelfAdjointView<Derived, UpLo>(derived()); }
  template <int UpLo> DenseSelfAdjointView<const Derived, UpLo> selfadjointView() const { return DenseSelfAdjointView<const Derived, UpLo>(derived()); }
  LLT<Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime> > llt() const;
  LDLT<Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime> > ldlt() const;
  Matrix<std::complex<Scalar>, RowsAtCompileTime, 1> eigenvalues() const;
  PartialPivLU<Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime> > lu() const;
  PartialPivLU<Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime> > partialPivLu() const;

  // ---- statics ----
  static PlainObject Zero() { PlainObject r; r.setZero(); return r; }
  static PlainObject Zero(Index n) { PlainObject r(n); r.setZero(); return r; }
  static PlainObject Zero(Index a, Index b) { PlainObject r(a, b); r.setZero(); return r; }
  static PlainObject Ones() { PlainObject r; r.setOnes(); return r; }
  static PlainObject Ones(Index n) { PlainObject r(n); r.setOnes(); return r; }
  static PlainObject Ones(Index a, Index b) { PlainObject r(a, b); r.setOnes(); return r; }
  static PlainObject Constant(const Scalar& s) { PlainObject r; r.setConstant(s); return r; }
  static PlainObject Constant(Index n, const Scalar& s) { PlainObject r(n); r.setConstant(s); return r; }
  static PlainObject Constant(Index a, Index b, const Scalar& s) { PlainObject r(a, b); r.setConstant(s); return r; }
  static PlainObject Identity() { PlainObject r; r.setIdentity(); return r; }
  static PlainObject Identity(Index a, Index b) { PlainObject r(a, b); r.setIdentity(); return r; }
  static PlainObject UnitX() { PlainObject r; r.setZero(); r[0] = 1; return r; }
  static PlainObject UnitY() { PlainObject r; r.setZero(); r[1] = 1; return r; }
  static PlainObject UnitZ() { PlainObject r; r.setZero(); r[2] = 1; return r; }
};

// comma initialiser: m << a, b, c;  (scalars and dense blocks, row by row)
template <class D> class CommaInitializer {
public:
  typedef typename traits<D>::Scalar Scalar;
  CommaInitializer(D& m, const Scalar& s) : m_(m), row_(0), col_(1), blockRows_(1) { m_.coeffRef(0, 0) = s; }
  template <class O> CommaInitializer(D& m, const MatrixBase<O>& o) : m_(m), row_(0), col_(o.cols()), blockRows_(o.rows()) { put(0, 0, o); }
  CommaInitializer& operator,(const Scalar& s) {
    if (col_ == m_.cols()) { row_ += blockRows_; col_ = 0; blockRows_ = 1; }
    m_.coeffRef(row_, col_++) = s;
    return *this;
  }
  template <class O> CommaInitializer& operator,(const MatrixBase<O>& o) {
    if (col_ == m_.cols()) { row_ += blockRows_; col_ = 0; blockRows_ = o.rows(); }
    put(row_, col_, o);
    col_ += o.cols();
    return *this;
  }
  D& finished() { return m_; }
private:
  template <class O> void put(Index r, Index c, const MatrixBase<O>& o) { for (Index j = 0; j < o.cols(); ++j) for (Index i = 0; i < o.rows(); ++i) m_.coeffRef(r + i, c + j) = o.coeff(i, j); }
  D& m_;
  Index row_, col_, blockRows_;
};
template <class D> CommaInitializer<D> operator<<(MatrixBase<D>& m, const typename traits<D>::Scalar& s) { return CommaInitializer<D>(m.derived(), s); }
template <class D, class O> CommaInitializer<D> operator<<(MatrixBase<D>& m, const MatrixBase<O>& o) { return CommaInitializer<D>(m.derived(), o); }
// temporaries (blocks): v.head<3>() << a, b, c;
template <class S, int R, int C> CommaInitializer<Block<S, R, C> > operator<<(Block<S, R, C>&& m, const typename std::remove_const<S>::type& s) { return CommaInitializer<Block<S, R, C> >(m, s); }

template <class D> std::ostream& operator<<(std::ostream& os, const MatrixBase<D>& m) {
  for (Index i = 0; i < m.rows(); ++i) {
    for (Index j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m.coeff(i, j);
    if (i + 1 < m.rows()) os << "\n";
  }
  return os;
}

// ------------------------------------------------------------------------------------------------------------------------
// Matrix
// ------------------------------------------------------------------------------------------------------------------------
template <class S, int R, int C, int Opt, int MR, int MC> struct traits<Matrix<S, R, C, Opt, MR, MC> > { typedef S Scalar; typedef S& Ref; enum { Rows = R, Cols = C }; };

namespace internal {
template <class S, int R, int C, bool Fixed = (R != Dynamic && C != Dynamic)> struct Storage;
template <class S, int R, int C> struct Storage<S, R, C, true> {
  S d[R * C > 0 ? R * C : 1];
  Storage() { for (int i = 0; i < R * C; ++i) d[i] = S(); }
  S* data() { return d; }
  const S* data() const { return d; }
  Index rows() const { return R; }
  Index cols() const { return C; }
  void resize(Index r, Index c) { assert(r == R && c == C); (void)r; (void)c; }
  void swap(Storage& o) { std::swap(*this, o); }
};
template <class S, int R, int C> struct Storage<S, R, C, false> {
  std::vector<S> d;
  Index r_, c_;
  Storage() : r_(R == Dynamic ? 0 : R), c_(C == Dynamic ? 0 : C) {}
  S* data() { return d.data(); }
  const S* data() const { return d.data(); }
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  void resize(Index r, Index c) { assert((R == Dynamic || r == R) && (C == Dynamic || c == C)); if (r * c != r_ * c_) { d.assign(size_t(r * c), S()); } r_ = r; c_ = c; }
  void swap(Storage& o) { d.swap(o.d); std::swap(r_, o.r_); std::swap(c_, o.c_); }
};
}  // namespace internal

template <class S, int R, int C, int Opt, int MR, int MC> class Matrix : public MatrixBase<Matrix<S, R, C, Opt, MR, MC> > {
public:
  typedef MatrixBase<Matrix> Base;
  typedef S Scalar;
  typedef Eigen::Index Index;
  typedef Map<Matrix, Unaligned> MapType;
  typedef const Map<const Matrix, Unaligned> ConstMapType;
  typedef Map<Matrix, Aligned> AlignedMapType;
  typedef const Map<const Matrix, Aligned> ConstAlignedMapType;
  enum { IsFixed = (R != Dynamic && C != Dynamic) };

  Matrix() {}
  Matrix(const Matrix& o) : Base(), st_(o.st_) {}
  Matrix& operator=(const Matrix& o) { st_ = o.st_; return *this; }
  template <class O> Matrix(const MatrixBase<O>& o) { this->assign(o); }
  template <class O> Matrix& operator=(const MatrixBase<O>& o) { return this->assign(o); }
  // (size) for dynamic vectors; for a fixed 1-vector the coefficient
  template <class T, class = typename std::enable_if<internal::is_arith<T>::value>::type> explicit Matrix(const T& n) {
    if (IsFixed) { if (R * C == 1) st_.data()[0] = S(n); else assert(Index(n) == R * C); }
    else if (C == 1) st_.resize(Index(n), 1);
    else if (R == 1) st_.resize(1, Index(n));
    else st_.resize(Index(n), Index(n));
  }
  explicit Matrix(const S* p) { for (Index i = 0; i < R * C; ++i) st_.data()[i] = p[i]; }
  // (rows, cols) for dynamic objects, (x, y) for fixed 2-vectors
  template <class T0, class T1, class = typename std::enable_if<internal::is_arith<T0>::value && internal::is_arith<T1>::value>::type> Matrix(const T0& a, const T1& b) {
    if (IsFixed && R * C == 2) { st_.data()[0] = S(a); st_.data()[1] = S(b); }
    else st_.resize(Index(a), Index(b));
  }
  Matrix(const S& a, const S& b, const S& c) { static_assert(R * C == 3, "3-vector"); S* d = st_.data(); d[0] = a; d[1] = b; d[2] = c; }
  Matrix(const S& a, const S& b, const S& c, const S& e) { static_assert(R * C == 4, "4-vector"); S* d = st_.data(); d[0] = a; d[1] = b; d[2] = c; d[3] = e; }

  S* data() { return st_.data(); }
  const S* data() const { return st_.data(); }
  Index rows() const { return st_.rows(); }
  Index cols() const { return st_.cols(); }
  Index outerStride() const { return st_.rows(); }
  void resize(Index r, Index c) { st_.resize(r, c); }
  void resize(Index n) { if (C == 1) st_.resize(n, 1); else if (R == 1) st_.resize(1, n); else { assert(R != Dynamic || C != Dynamic); st_.resize(R == Dynamic ? n / C : R, C == Dynamic ? n / R : C); } }
  void resizeLike(Index r, Index c) { if (r != rows() || c != cols()) st_.resize(r, c); }
  void conservativeResize(Index r, Index c) {
    Matrix t(r, c);
    for (Index j = 0; j < std::min(c, cols()); ++j) for (Index i = 0; i < std::min(r, rows()); ++i) t.coeffRef(i, j) = this->coeff(i, j);
    st_.swap(t.st_);
  }
  void conservativeResize(Index n) { if (C == 1) conservativeResize(n, 1); else conservativeResize(1, n); }
  void swap(Matrix& o) { st_.swap(o.st_); }
  template <class O> void swap(MatrixBase<O>& o) { Base::swap(o); }
  static MapType Map(S* p) { return MapType(p); }
  static MapType Map(S* p, Index n) { return MapType(p, n); }
  static MapType Map(S* p, Index r, Index c) { return MapType(p, r, c); }
  static ConstMapType Map(const S* p) { return ConstMapType(p); }
  static ConstMapType Map(const S* p, Index n) { return ConstMapType(p, n); }
  static ConstMapType Map(const S* p, Index r, Index c) { return ConstMapType(p, r, c); }
  static AlignedMapType MapAligned(S* p) { return AlignedMapType(p); }
  static AlignedMapType MapAligned(S* p, Index n) { return AlignedMapType(p, n); }
  static AlignedMapType MapAligned(S* p, Index r, Index c) { return AlignedMapType(p, r, c); }
  static Matrix Random() { Matrix m; for (Index i = 0; i < m.size(); ++i) m.data()[i] = S(2.0 * std::rand() / RAND_MAX - 1.0); return m; }
  static Matrix Random(Index r, Index c) { Matrix m(r, c); for (Index i = 0; i < m.size(); ++i) m.data()[i] = S(2.0 * std::rand() / RAND_MAX - 1.0); return m; }
private:
  internal::Storage<S, R, C> st_;
};

#define EIGENMINI_TYPEDEFS(T, s) \
  typedef Matrix<T, 2, 2> Matrix2##s; typedef Matrix<T, 3, 3> Matrix3##s; typedef Matrix<T, 4, 4> Matrix4##s; typedef Matrix<T, Dynamic, Dynamic> MatrixX##s; \
  typedef Matrix<T, 2, 1> Vector2##s; typedef Matrix<T, 3, 1> Vector3##s; typedef Matrix<T, 4, 1> Vector4##s; typedef Matrix<T, Dynamic, 1> VectorX##s; \
  typedef Matrix<T, 1, 2> RowVector2##s; typedef Matrix<T, 1, 3> RowVector3##s; typedef Matrix<T, 1, 4> RowVector4##s; typedef Matrix<T, 1, Dynamic> RowVectorX##s;
EIGENMINI_TYPEDEFS(double, d)
EIGENMINI_TYPEDEFS(float, f)
EIGENMINI_TYPEDEFS(int, i)
typedef Matrix<std::complex<double>, Dynamic, 1> VectorXcd;

// ------------------------------------------------------------------------------------------------------------------------
// Block and Map: non-owning strided views
// ------------------------------------------------------------------------------------------------------------------------
template <class S, int R, int C> struct traits<Block<S, R, C> > { typedef typename std::remove_const<S>::type Scalar; typedef S& Ref; enum { Rows = R, Cols = C }; };
template <class S, int R, int C> class Block : public MatrixBase<Block<S, R, C> > {
public:
  typedef MatrixBase<Block> Base;
  typedef typename std::remove_const<S>::type Scalar;
  Block(S* p, Index r, Index c, Index stride) : p_(p), r_(r), c_(c), s_(stride) {}
  Block(const Block& o) : Base(), p_(o.p_), r_(o.r_), c_(o.c_), s_(o.s_) {}
  Block& operator=(const Block& o) { return this->assign(o); }
  template <class O> Block& operator=(const MatrixBase<O>& o) { return this->assign(o); }
  S* data() const { return p_; }
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Index outerStride() const { return s_; }
  void resizeLike(Index r, Index c) const { assert(r == r_ && c == c_); (void)r; (void)c; }
  void resize(Index r, Index c) const { resizeLike(r, c); }
  void resize(Index n) const { assert(n == r_ * c_); (void)n; }
private:
  S* p_;
  Index r_, c_, s_;
};

template <class M, int MapOpt, class Stride> struct traits<Map<M, MapOpt, Stride> > {
  typedef typename traits<typename std::remove_const<M>::type>::Scalar Scalar;
  typedef typename std::conditional<std::is_const<M>::value, const Scalar&, Scalar&>::type Ref;
  enum { Rows = traits<typename std::remove_const<M>::type>::Rows, Cols = traits<typename std::remove_const<M>::type>::Cols };
};
template <class M, int MapOpt, class Stride> class Map : public MatrixBase<Map<M, MapOpt, Stride> > {
public:
  typedef MatrixBase<Map> Base;
  typedef typename traits<Map>::Scalar Scalar;
  typedef typename std::conditional<std::is_const<M>::value, const Scalar, Scalar>::type S;
  enum { R = traits<Map>::Rows, C = traits<Map>::Cols };
  explicit Map(S* p) : p_(p), r_(R), c_(C) { assert(R != Dynamic && C != Dynamic); }
  Map(S* p, Index n) : p_(p), r_(C == 1 ? n : (R == Dynamic ? n : R)), c_(C == 1 ? 1 : (R == 1 ? n : C)) {}
  Map(S* p, Index r, Index c) : p_(p), r_(r), c_(c) {}
  Map(const Map& o) : Base(), p_(o.p_), r_(o.r_), c_(o.c_) {}
  Map& operator=(const Map& o) { return this->assign(o); }
  template <class O> Map& operator=(const MatrixBase<O>& o) { return this->assign(o); }
  S* data() const { return p_; }
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Index outerStride() const { return r_; }
  void resizeLike(Index r, Index c) const { assert(r == r_ && c == c_); (void)r; (void)c; }
  void resize(Index r, Index c) const { resizeLike(r, c); }
  void resize(Index n) const { assert(n == r_ * c_); (void)n; }
private:
  S* p_;
  Index r_, c_;
};

template <class D> class NoAlias {
public:
  explicit NoAlias(D& d) : d_(d) {}
  template <class O> D& operator=(const MatrixBase<O>& o) { return d_.assign(o); }
  template <class O> D& operator+=(const MatrixBase<O>& o) { return d_ += o; }
  template <class O> D& operator-=(const MatrixBase<O>& o) { return d_ -= o; }
private:
  D& d_;
};

// m.array() on an lvalue: only coefficient-wise compound assignment with a scalar or another array is needed
template <class D> class ArrayWrapper {
public:
  typedef typename traits<D>::Scalar Scalar;
  explicit ArrayWrapper(D& d) : d_(d) {}
  ArrayWrapper& operator+=(const Scalar& s) { for (Index j = 0; j < d_.cols(); ++j) for (Index i = 0; i < d_.rows(); ++i) d_.coeffRef(i, j) += s; return *this; }
  ArrayWrapper& operator-=(const Scalar& s) { for (Index j = 0; j < d_.cols(); ++j) for (Index i = 0; i < d_.rows(); ++i) d_.coeffRef(i, j) -= s; return *this; }
  ArrayWrapper& operator*=(const Scalar& s) { for (Index j = 0; j < d_.cols(); ++j) for (Index i = 0; i < d_.rows(); ++i) d_.coeffRef(i, j) *= s; return *this; }
  ArrayWrapper& operator/=(const Scalar& s) { for (Index j = 0; j < d_.cols(); ++j) for (Index i = 0; i < d_.rows(); ++i) d_.coeffRef(i, j) /= s; return *this; }
private:
  D& d_;
};

template <class Derived> class MatrixBase<Derived>::DiagonalProxy {
public:
  typedef typename traits<Derived>::Scalar Scalar;
  typedef Matrix<Scalar, internal::pick<traits<Derived>::Rows, traits<Derived>::Cols>::value, 1> Vec;
  explicit DiagonalProxy(Derived& d) : d_(d) {}
  Index size() const { return std::min(d_.rows(), d_.cols()); }
  Scalar& operator()(Index i) { return d_.coeffRef(i, i); }
  Scalar& operator[](Index i) { return d_.coeffRef(i, i); }
  const Scalar& coeff(Index i) const { return d_.coeff(i, i); }
  DiagonalProxy& array() { return *this; }
  DiagonalProxy& operator+=(const Scalar& s) { for (Index i = 0; i < size(); ++i) d_.coeffRef(i, i) += s; return *this; }
  DiagonalProxy& operator-=(const Scalar& s) { for (Index i = 0; i < size(); ++i) d_.coeffRef(i, i) -= s; return *this; }
  DiagonalProxy& operator*=(const Scalar& s) { for (Index i = 0; i < size(); ++i) d_.coeffRef(i, i) *= s; return *this; }
  template <class O> DiagonalProxy& operator=(const MatrixBase<O>& o) { for (Index i = 0; i < size(); ++i) d_.coeffRef(i, i) = o.coeff(i); return *this; }
  template <class O> DiagonalProxy& operator+=(const MatrixBase<O>& o) { for (Index i = 0; i < size(); ++i) d_.coeffRef(i, i) += o.coeff(i); return *this; }
  template <class O> DiagonalProxy& operator-=(const MatrixBase<O>& o) { for (Index i = 0; i < size(); ++i) d_.coeffRef(i, i) -= o.coeff(i); return *this; }
  DiagonalProxy& setConstant(const Scalar& s) { for (Index i = 0; i < size(); ++i) d_.coeffRef(i, i) = s; return *this; }
  Vec eval() const { Vec v(size()); for (Index i = 0; i < size(); ++i) v[i] = coeff(i); return v; }
  operator Vec() const { return eval(); }
  Scalar maxCoeff() const { return eval().maxCoeff(); }
  Scalar minCoeff() const { return eval().minCoeff(); }
  Scalar sum() const { return eval().sum(); }
private:
  Derived& d_;
};

// ------------------------------------------------------------------------------------------------------------------------
// arithmetic operators (eager)
// ------------------------------------------------------------------------------------------------------------------------
#define EIGENMINI_BIN_RESULT(A, B) Matrix<typename traits<A>::Scalar, internal::pick<traits<A>::Rows, traits<B>::Rows>::value, internal::pick<traits<A>::Cols, traits<B>::Cols>::value>
template <class A, class B> EIGENMINI_BIN_RESULT(A, B) operator+(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  assert(a.rows() == b.rows() && a.cols() == b.cols());
  EIGENMINI_BIN_RESULT(A, B) r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.coeffRef(i, j) = a.coeff(i, j) + b.coeff(i, j);
  return r;
}
template <class A, class B> EIGENMINI_BIN_RESULT(A, B) operator-(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  assert(a.rows() == b.rows() && a.cols() == b.cols());
  EIGENMINI_BIN_RESULT(A, B) r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.coeffRef(i, j) = a.coeff(i, j) - b.coeff(i, j);
  return r;
}
template <class A> typename MatrixBase<A>::PlainObject operator*(const MatrixBase<A>& a, const typename traits<A>::Scalar& s) {
  typename MatrixBase<A>::PlainObject r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.coeffRef(i, j) = a.coeff(i, j) * s;
  return r;
}
template <class A> typename MatrixBase<A>::PlainObject operator*(const typename traits<A>::Scalar& s, const MatrixBase<A>& a) {
  typename MatrixBase<A>::PlainObject r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.coeffRef(i, j) = s * a.coeff(i, j);
  return r;
}
template <class A> typename MatrixBase<A>::PlainObject operator/(const MatrixBase<A>& a, const typename traits<A>::Scalar& s) {
  typename MatrixBase<A>::PlainObject r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r.coeffRef(i, j) = a.coeff(i, j) / s;
  return r;
}
// matrix product: coefficient (i,j) = ((a_i0 b_0j + a_i1 b_1j) + a_i2 b_2j) + ...
template <class A, class B> Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> operator*(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  assert(a.cols() == b.rows());
  Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> r(a.rows(), b.cols());
  const Index K = a.cols();
  for (Index j = 0; j < b.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) {
      typename traits<A>::Scalar s = K ? a.coeff(i, 0) * b.coeff(0, j) : typename traits<A>::Scalar(0);
      for (Index k = 1; k < K; ++k) s += a.coeff(i, k) * b.coeff(k, j);
      r.coeffRef(i, j) = s;
    }
  return r;
}

// ------------------------------------------------------------------------------------------------------------------------
// determinant / inverse
// ------------------------------------------------------------------------------------------------------------------------
namespace internal {
template <class M> typename M::Scalar det3(const M& m, int a, int b, int c) {  // cofactor expansion helper on columns 0..2 with rows a,b,c
  return m.coeff(a, 0) * (m.coeff(b, 1) * m.coeff(c, 2) - m.coeff(c, 1) * m.coeff(b, 2));
}
template <class M> typename M::Scalar cofactor3(const M& m, int i, int j) {
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m.coeff(i1, j1) * m.coeff(i2, j2) - m.coeff(i1, j2) * m.coeff(i2, j1);
}
// partial-pivot LU of a dynamic copy; returns the sign of the permutation (0 if singular), perm applied to rows
template <class S> int lu_inplace(Matrix<S, Dynamic, Dynamic>& a, std::vector<Index>& piv) {
  const Index n = a.rows();
  piv.resize(size_t(n));
  int sign = 1;
  for (Index k = 0; k < n; ++k) {
    Index p = k;
    S best = std::abs(a.coeff(k, k));
    for (Index i = k + 1; i < n; ++i) if (std::abs(a.coeff(i, k)) > best) { best = std::abs(a.coeff(i, k)); p = i; }
    piv[size_t(k)] = p;
    if (best == S(0)) { sign = 0; continue; }
    if (p != k) { for (Index j = 0; j < n; ++j) std::swap(a.coeffRef(k, j), a.coeffRef(p, j)); sign = -sign; }
    for (Index i = k + 1; i < n; ++i) a.coeffRef(i, k) /= a.coeff(k, k);
    for (Index j = k + 1; j < n; ++j) for (Index i = k + 1; i < n; ++i) a.coeffRef(i, j) -= a.coeff(i, k) * a.coeff(k, j);
  }
  return sign;
}
}  // namespace internal

template <class D> typename MatrixBase<D>::Scalar MatrixBase<D>::determinant() const {
  const Index n = rows();
  assert(n == cols());
  if (n == 0) return Scalar(1);
  if (n == 1) return coeff(0, 0);
  if (n == 2) return coeff(0, 0) * coeff(1, 1) - coeff(1, 0) * coeff(0, 1);
  if (n == 3) return internal::det3(*this, 0, 1, 2) - internal::det3(*this, 1, 0, 2) + internal::det3(*this, 2, 0, 1);
  Matrix<Scalar, Dynamic, Dynamic> a(*this);
  std::vector<Index> piv;
  int sign = internal::lu_inplace(a, piv);
  Scalar d = Scalar(sign);
  for (Index i = 0; i < n; ++i) d *= a.coeff(i, i);
  return d;
}

template <class D> typename MatrixBase<D>::PlainObject MatrixBase<D>::inverse() const {
  const Index n = rows();
  assert(n == cols());
  PlainObject r(n, n);
  if (n == 1) { r.coeffRef(0, 0) = Scalar(1) / coeff(0, 0); return r; }
  if (n == 2) {
    const Scalar invdet = Scalar(1) / determinant();
    r.coeffRef(0, 0) = coeff(1, 1) * invdet; r.coeffRef(1, 0) = -coeff(1, 0) * invdet;
    r.coeffRef(0, 1) = -coeff(0, 1) * invdet; r.coeffRef(1, 1) = coeff(0, 0) * invdet;
    return r;
  }
  if (n == 3) {  // Eigen: cofactors of the first column give the determinant, result = cofactor matrix transposed * (1/det)
    const Scalar c00 = internal::cofactor3(*this, 0, 0), c10 = internal::cofactor3(*this, 1, 0), c20 = internal::cofactor3(*this, 2, 0);
    const Scalar det = (c00 * coeff(0, 0) + c10 * coeff(1, 0)) + c20 * coeff(2, 0);
    const Scalar invdet = Scalar(1) / det;
    r.coeffRef(0, 0) = c00 * invdet; r.coeffRef(0, 1) = c10 * invdet; r.coeffRef(0, 2) = c20 * invdet;
    r.coeffRef(1, 0) = internal::cofactor3(*this, 0, 1) * invdet; r.coeffRef(1, 1) = internal::cofactor3(*this, 1, 1) * invdet; r.coeffRef(1, 2) = internal::cofactor3(*this, 2, 1) * invdet;
    r.coeffRef(2, 0) = internal::cofactor3(*this, 0, 2) * invdet; r.coeffRef(2, 1) = internal::cofactor3(*this, 1, 2) * invdet; r.coeffRef(2, 2) = internal::cofactor3(*this, 2, 2) * invdet;
    return r;
  }
  // general: solve A X = I with partial-pivot LU (Eigen: PartialPivLU)
  Matrix<Scalar, Dynamic, Dynamic> a(*this);
  std::vector<Index> piv;
  internal::lu_inplace(a, piv);
  Matrix<Scalar, Dynamic, Dynamic> x(n, n);
  x.setIdentity();
  for (Index k = 0; k < n; ++k) if (piv[size_t(k)] != k) for (Index j = 0; j < n; ++j) std::swap(x.coeffRef(k, j), x.coeffRef(piv[size_t(k)], j));
  for (Index j = 0; j < n; ++j) {
    for (Index i = 0; i < n; ++i) { Scalar s = x.coeff(i, j); for (Index k = 0; k < i; ++k) s -= a.coeff(i, k) * x.coeff(k, j); x.coeffRef(i, j) = s; }
    for (Index i = n - 1; i >= 0; --i) { Scalar s = x.coeff(i, j); for (Index k = i + 1; k < n; ++k) s -= a.coeff(i, k) * x.coeff(k, j); x.coeffRef(i, j) = s / a.coeff(i, i); }
  }
  r.assign(x);
  return r;
}

// ------------------------------------------------------------------------------------------------------------------------
// dense Cholesky
// ------------------------------------------------------------------------------------------------------------------------
template <class M> class LLT {
public:
  typedef typename M::Scalar Scalar;
  LLT() : ok_(false) {}
  template <class O> explicit LLT(const MatrixBase<O>& a) { compute(a); }
  template <class O> LLT& compute(const MatrixBase<O>& a) {
    const Index n = a.rows();
    l_.resize(n, n);
    l_.setZero();
    ok_ = true;
    for (Index j = 0; j < n; ++j) {
      Scalar d = a.coeff(j, j);
      for (Index k = 0; k < j; ++k) d -= l_.coeff(j, k) * l_.coeff(j, k);
      if (!(d > Scalar(0))) { ok_ = false; return *this; }
      d = std::sqrt(d);
      l_.coeffRef(j, j) = d;
      for (Index i = j + 1; i < n; ++i) {
        Scalar s = a.coeff(i, j);
        for (Index k = 0; k < j; ++k) s -= l_.coeff(i, k) * l_.coeff(j, k);
        l_.coeffRef(i, j) = s / d;
      }
    }
    return *this;
  }
  template <class B> typename MatrixBase<B>::PlainObject solve(const MatrixBase<B>& b) const {
    typename MatrixBase<B>::PlainObject x(b);
    const Index n = l_.rows();
    for (Index c = 0; c < x.cols(); ++c) {
      for (Index i = 0; i < n; ++i) { Scalar s = x.coeff(i, c); for (Index k = 0; k < i; ++k) s -= l_.coeff(i, k) * x.coeff(k, c); x.coeffRef(i, c) = s / l_.coeff(i, i); }
      for (Index i = n - 1; i >= 0; --i) { Scalar s = x.coeff(i, c); for (Index k = i + 1; k < n; ++k) s -= l_.coeff(k, i) * x.coeff(k, c); x.coeffRef(i, c) = s / l_.coeff(i, i); }
    }
    return x;
  }
  Matrix<Scalar, Dynamic, Dynamic> matrixL() const { return l_; }
  Matrix<Scalar, Dynamic, Dynamic> matrixU() const { return l_.transpose(); }
  ComputationInfo info() const { return ok_ ? Success : NumericalIssue; }
private:
  Matrix<Scalar, Dynamic, Dynamic> l_;
  bool ok_;
};

// LDLT with diagonal pivoting (Eigen's LDLT: at step k the largest remaining |diagonal| is brought to position k)
template <class M> class LDLT {
public:
  typedef typename M::Scalar Scalar;
  LDLT() : sign_(0), ok_(false) {}
  template <class O> explicit LDLT(const MatrixBase<O>& a) { compute(a); }
  template <class O> LDLT& compute(const MatrixBase<O>& a) {
    const Index n = a.rows();
    m_.assign(a);
    perm_.resize(size_t(n));
    for (Index i = 0; i < n; ++i) perm_[size_t(i)] = i;
    bool pos = true, neg = true;
    ok_ = true;
    for (Index k = 0; k < n; ++k) {
      Index p = k;
      Scalar best = std::abs(m_.coeff(k, k));
      for (Index i = k + 1; i < n; ++i) if (std::abs(m_.coeff(i, i)) > best) { best = std::abs(m_.coeff(i, i)); p = i; }
      trans_.push_back(p);
      if (p != k) {  // symmetric row/column swap on the lower triangle + the already computed part of L
        for (Index j = 0; j < n; ++j) std::swap(m_.coeffRef(k, j), m_.coeffRef(p, j));
        for (Index i = 0; i < n; ++i) std::swap(m_.coeffRef(i, k), m_.coeffRef(i, p));
        std::swap(perm_[size_t(k)], perm_[size_t(p)]);
      }
      // d_k = a_kk - sum_j l_kj^2 d_j ; the trailing part is updated eagerly (right-looking), so m_(k,k) already holds d_k
      const Scalar d = m_.coeff(k, k);
      if (d > Scalar(0)) neg = false; else if (d < Scalar(0)) pos = false; else { pos = pos && true; neg = neg && true; }
      if (d == Scalar(0)) { for (Index i = k + 1; i < n; ++i) m_.coeffRef(i, k) = Scalar(0); continue; }
      for (Index i = k + 1; i < n; ++i) m_.coeffRef(i, k) /= d;
      for (Index j = k + 1; j < n; ++j) {
        const Scalar ljd = m_.coeff(j, k) * d;
        for (Index i = j; i < n; ++i) m_.coeffRef(i, j) -= m_.coeff(i, k) * ljd;
        for (Index i = k + 1; i < j; ++i) m_.coeffRef(i, j) = m_.coeff(j, i);  // keep the square symmetric for later swaps
      }
    }
    sign_ = pos ? 1 : (neg ? -1 : 0);
    return *this;
  }
  bool isPositive() const { return sign_ == 1; }
  bool isNegative() const { return sign_ == -1; }
  ComputationInfo info() const { return ok_ ? Success : NumericalIssue; }
  Matrix<Scalar, Dynamic, 1> vectorD() const { Matrix<Scalar, Dynamic, 1> d(m_.rows()); for (Index i = 0; i < m_.rows(); ++i) d[i] = m_.coeff(i, i); return d; }
  template <class B> typename MatrixBase<B>::PlainObject solve(const MatrixBase<B>& b) const {
    const Index n = m_.rows();
    typename MatrixBase<B>::PlainObject x(b.rows(), b.cols());
    for (Index c = 0; c < b.cols(); ++c) {
      std::vector<Scalar> y(size_t(n), Scalar(0));
      for (Index i = 0; i < n; ++i) y[size_t(i)] = b.coeff(perm_[size_t(i)], c);
      for (Index i = 0; i < n; ++i) { Scalar s = y[size_t(i)]; for (Index k = 0; k < i; ++k) s -= m_.coeff(i, k) * y[size_t(k)]; y[size_t(i)] = s; }
      for (Index i = 0; i < n; ++i) { const Scalar d = m_.coeff(i, i); y[size_t(i)] = (std::abs(d) > std::numeric_limits<Scalar>::min()) ? y[size_t(i)] / d : Scalar(0); }
      for (Index i = n - 1; i >= 0; --i) { Scalar s = y[size_t(i)]; for (Index k = i + 1; k < n; ++k) s -= m_.coeff(k, i) * y[size_t(k)]; y[size_t(i)] = s; }
      for (Index i = 0; i < n; ++i) x.coeffRef(perm_[size_t(i)], c) = y[size_t(i)];
    }
    return x;
  }
private:
  Matrix<Scalar, Dynamic, Dynamic> m_;
  std::vector<Index> perm_, trans_;
  int sign_;
  bool ok_;
};
template <class D> LLT<Matrix<typename MatrixBase<D>::Scalar, MatrixBase<D>::RowsAtCompileTime, MatrixBase<D>::ColsAtCompileTime> > MatrixBase<D>::llt() const {
  return LLT<Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime> >(*this);
}
template <class D> LDLT<Matrix<typename MatrixBase<D>::Scalar, MatrixBase<D>::RowsAtCompileTime, MatrixBase<D>::ColsAtCompileTime> > MatrixBase<D>::ldlt() const {
  return LDLT<Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime> >(*this);
}

// partial-pivot LU (Eigen's PartialPivLU: the default behind .lu())
template <class M> class PartialPivLU {
public:
  typedef typename M::Scalar Scalar;
  PartialPivLU() : sign_(0) {}
  template <class O> explicit PartialPivLU(const MatrixBase<O>& a) { compute(a); }
  template <class O> PartialPivLU& compute(const MatrixBase<O>& a) { lu_.assign(a); sign_ = internal::lu_inplace(lu_, piv_); return *this; }
  template <class B> typename MatrixBase<B>::PlainObject solve(const MatrixBase<B>& b) const {
    typename MatrixBase<B>::PlainObject x(b);
    const Index n = lu_.rows();
    for (Index k = 0; k < n; ++k) if (piv_[size_t(k)] != k) for (Index j = 0; j < x.cols(); ++j) std::swap(x.coeffRef(k, j), x.coeffRef(piv_[size_t(k)], j));
    for (Index j = 0; j < x.cols(); ++j) {
      for (Index i = 0; i < n; ++i) { Scalar s = x.coeff(i, j); for (Index k = 0; k < i; ++k) s -= lu_.coeff(i, k) * x.coeff(k, j); x.coeffRef(i, j) = s; }
      for (Index i = n - 1; i >= 0; --i) { Scalar s = x.coeff(i, j); for (Index k = i + 1; k < n; ++k) s -= lu_.coeff(i, k) * x.coeff(k, j); x.coeffRef(i, j) = s / lu_.coeff(i, i); }
    }
    return x;
  }
  Scalar determinant() const { Scalar d = Scalar(sign_); for (Index i = 0; i < lu_.rows(); ++i) d *= lu_.coeff(i, i); return d; }
  M inverse() const { M id(lu_.rows(), lu_.cols()); id.setIdentity(); return solve(id); }
private:
  Matrix<Scalar, Dynamic, Dynamic> lu_;
  std::vector<Index> piv_;
  int sign_;
};
template <class D> PartialPivLU<Matrix<typename MatrixBase<D>::Scalar, MatrixBase<D>::RowsAtCompileTime, MatrixBase<D>::ColsAtCompileTime> > MatrixBase<D>::lu() const {
  return PartialPivLU<Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime> >(*this);
}
template <class D> PartialPivLU<Matrix<typename MatrixBase<D>::Scalar, MatrixBase<D>::RowsAtCompileTime, MatrixBase<D>::ColsAtCompileTime> > MatrixBase<D>::partialPivLu() const {
  return PartialPivLU<Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime> >(*this);
}

// dense selfadjoint view: A.selfadjointView<Upper>() used as an rvalue gives the full symmetric matrix
template <class D, int UpLo> class DenseSelfAdjointView {
public:
  typedef typename traits<typename std::remove_const<D>::type>::Scalar Scalar;
  typedef Matrix<Scalar, traits<typename std::remove_const<D>::type>::Rows, traits<typename std::remove_const<D>::type>::Cols> Plain;
  explicit DenseSelfAdjointView(D& d) : d_(d) {}
  Plain full() const {
    Plain r(d_.rows(), d_.cols());
    for (Index j = 0; j < d_.cols(); ++j) for (Index i = 0; i < d_.rows(); ++i) {
      const bool stored = (UpLo == Upper) ? (i <= j) : (i >= j);
      r.coeffRef(i, j) = stored ? d_.coeff(i, j) : d_.coeff(j, i);
    }
    return r;
  }
  operator Plain() const { return full(); }
  template <class O> typename MatrixBase<O>::PlainObject operator*(const MatrixBase<O>& o) const { return full() * o; }
  LLT<Plain> llt() const { return LLT<Plain>(full()); }
  LDLT<Plain> ldlt() const { return LDLT<Plain>(full()); }
private:
  D& d_;
};

// symmetric eigenvalues (cyclic Jacobi); only what verifyInformationMatrices-style checks need
template <class M> class SelfAdjointEigenSolver {
public:
  typedef typename M::Scalar Scalar;
  typedef Matrix<Scalar, traits<M>::Rows, 1> RealVectorType;
  SelfAdjointEigenSolver() {}
  template <class O> explicit SelfAdjointEigenSolver(const MatrixBase<O>& a, int = ComputeEigenvectors) { compute(a); }
  template <class O> SelfAdjointEigenSolver& compute(const MatrixBase<O>& a0, int = ComputeEigenvectors) {
    const Index n = a0.rows();
    Matrix<Scalar, Dynamic, Dynamic> a(a0), v(n, n);
    v.setIdentity();
    for (int sweep = 0; sweep < 64; ++sweep) {
      Scalar off = 0;
      for (Index p = 0; p < n; ++p) for (Index q = p + 1; q < n; ++q) off += a.coeff(p, q) * a.coeff(p, q);
      if (off < std::numeric_limits<Scalar>::min()) break;
      for (Index p = 0; p < n; ++p) for (Index q = p + 1; q < n; ++q) {
        if (a.coeff(p, q) == Scalar(0)) continue;
        const Scalar theta = (a.coeff(q, q) - a.coeff(p, p)) / (2 * a.coeff(p, q));
        const Scalar t = (theta >= 0 ? 1 : -1) / (std::abs(theta) + std::sqrt(theta * theta + 1));
        const Scalar c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (Index k = 0; k < n; ++k) { const Scalar akp = a.coeff(k, p), akq = a.coeff(k, q); a.coeffRef(k, p) = c * akp - s * akq; a.coeffRef(k, q) = s * akp + c * akq; }
        for (Index k = 0; k < n; ++k) { const Scalar apk = a.coeff(p, k), aqk = a.coeff(q, k); a.coeffRef(p, k) = c * apk - s * aqk; a.coeffRef(q, k) = s * apk + c * aqk; }
        for (Index k = 0; k < n; ++k) { const Scalar vkp = v.coeff(k, p), vkq = v.coeff(k, q); v.coeffRef(k, p) = c * vkp - s * vkq; v.coeffRef(k, q) = s * vkp + c * vkq; }
      }
    }
    std::vector<Index> order(static_cast<size_t>(n));
    for (Index i = 0; i < n; ++i) order[size_t(i)] = i;
    std::sort(order.begin(), order.end(), [&](Index x, Index y) { return a.coeff(x, x) < a.coeff(y, y); });
    vals_.resize(n);
    vecs_.resize(n, n);
    for (Index i = 0; i < n; ++i) { vals_[i] = a.coeff(order[size_t(i)], order[size_t(i)]); for (Index k = 0; k < n; ++k) vecs_.coeffRef(k, i) = v.coeff(k, order[size_t(i)]); }
    return *this;
  }
  const RealVectorType& eigenvalues() const { return vals_; }
  const Matrix<Scalar, Dynamic, Dynamic>& eigenvectors() const { return vecs_; }
  ComputationInfo info() const { return Success; }
private:
  RealVectorType vals_;
  Matrix<Scalar, Dynamic, Dynamic> vecs_;
};
template <class D> Matrix<std::complex<typename MatrixBase<D>::Scalar>, MatrixBase<D>::RowsAtCompileTime, 1> MatrixBase<D>::eigenvalues() const {
  SelfAdjointEigenSolver<PlainObject> es(*this);  // used by the reference on symmetric matrices only
  Matrix<std::complex<Scalar>, RowsAtCompileTime, 1> r(rows());
  for (Index i = 0; i < rows(); ++i) r.data()[i] = std::complex<Scalar>(es.eigenvalues()[i], 0);
  return r;
}

// ------------------------------------------------------------------------------------------------------------------------
// Geometry: Quaternion, AngleAxis, Transform
// ------------------------------------------------------------------------------------------------------------------------
template <class S> class Quaternion {
public:
  typedef S Scalar;
  typedef Matrix<S, 4, 1> Coefficients;
  typedef Matrix<S, 3, 1> Vector3;
  typedef Matrix<S, 3, 3> Matrix3;
  Quaternion() {}
  Quaternion(const S& w, const S& x, const S& y, const S& z) { c_[0] = x; c_[1] = y; c_[2] = z; c_[3] = w; }
  explicit Quaternion(const S* p) { for (int i = 0; i < 4; ++i) c_[i] = p[i]; }
  template <class D> explicit Quaternion(const MatrixBase<D>& m) {
    if (m.rows() == 3 && m.cols() == 3) fromRotationMatrix(m);
    else { assert(m.size() == 4); for (int i = 0; i < 4; ++i) c_[i] = m.coeff(i); }
  }
  explicit Quaternion(const AngleAxis<S>& aa);
  S& x() { return c_[0]; } S& y() { return c_[1]; } S& z() { return c_[2]; } S& w() { return c_[3]; }
  const S& x() const { return c_[0]; } const S& y() const { return c_[1]; } const S& z() const { return c_[2]; } const S& w() const { return c_[3]; }
  Coefficients& coeffs() { return c_; }
  const Coefficients& coeffs() const { return c_; }
  Block<S, 3, 1> vec() { return c_.template head<3>(); }
  Block<const S, 3, 1> vec() const { return c_.template head<3>(); }
  Quaternion& setIdentity() { c_[0] = c_[1] = c_[2] = S(0); c_[3] = S(1); return *this; }
  static Quaternion Identity() { return Quaternion(S(1), S(0), S(0), S(0)); }
  S squaredNorm() const { return c_.squaredNorm(); }
  S norm() const { return c_.norm(); }
  void normalize() { c_.normalize(); }
  Quaternion normalized() const { Quaternion q(*this); q.normalize(); return q; }
  Quaternion conjugate() const { return Quaternion(c_[3], -c_[0], -c_[1], -c_[2]); }
  Quaternion inverse() const {
    const S n2 = squaredNorm();
    if (n2 > S(0)) { Quaternion q = conjugate(); q.c_ /= n2; return q; }
    Quaternion q; q.c_.setZero(); return q;
  }
  S dot(const Quaternion& o) const { return c_.dot(o.c_); }
  template <class T> Quaternion<T> cast() const { return Quaternion<T>(T(w()), T(x()), T(y()), T(z())); }
  Quaternion operator*(const Quaternion& b) const {
    const Quaternion& a = *this;
    return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                      a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                      a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                      a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
  }
  Quaternion& operator*=(const Quaternion& b) { *this = (*this) * b; return *this; }
  // Eigen's _transformVector: uv = 2 (q.vec x v); v + w uv + q.vec x uv
  template <class D> Vector3 operator*(const MatrixBase<D>& v) const { return _transformVector(Vector3(v)); }
  Vector3 _transformVector(const Vector3& v) const {
    Vector3 qv(c_[0], c_[1], c_[2]);
    Vector3 uv = qv.cross(v);
    uv += uv;
    return v + c_[3] * uv + qv.cross(uv);
  }
  Matrix3 toRotationMatrix() const {
    Matrix3 r;
    const S tx = S(2) * x(), ty = S(2) * y(), tz = S(2) * z();
    const S twx = tx * w(), twy = ty * w(), twz = tz * w();
    const S txx = tx * x(), txy = ty * x(), txz = tz * x();
    const S tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
    r(0, 0) = S(1) - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
    r(1, 0) = txy + twz; r(1, 1) = S(1) - (txx + tzz); r(1, 2) = tyz - twx;
    r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = S(1) - (txx + tyy);
    return r;
  }
  Matrix3 matrix() const { return toRotationMatrix(); }
  template <class D> Quaternion& operator=(const MatrixBase<D>& m) { fromRotationMatrix(m); return *this; }
  Quaternion& operator=(const AngleAxis<S>& aa) { *this = Quaternion(aa); return *this; }
  Quaternion slerp(const S& t, const Quaternion& o) const {
    const S one = S(1) - std::numeric_limits<S>::epsilon();
    const S d = dot(o), ad = std::abs(d);
    S s0, s1;
    if (ad >= one) { s0 = S(1) - t; s1 = t; }
    else { const S th = std::acos(ad), st = std::sin(th); s0 = std::sin((S(1) - t) * th) / st; s1 = std::sin(t * th) / st; }
    if (d < S(0)) s1 = -s1;
    Quaternion q; q.c_ = s0 * c_ + s1 * o.c_; return q;
  }
  S angularDistance(const Quaternion& o) const { const Quaternion d = (*this) * o.conjugate(); return S(2) * std::atan2(d.vec().norm(), std::abs(d.w())); }
private:
  // Shoemake, "Quaternion Calculus and Fast Animation" (what Eigen's quaternionbase_assign_impl<Other,3,3> does)
  template <class D> void fromRotationMatrix(const MatrixBase<D>& m) {
    S t = m.coeff(0, 0) + m.coeff(1, 1) + m.coeff(2, 2);
    if (t > S(0)) {
      t = std::sqrt(t + S(1.0));
      w() = S(0.5) * t;
      t = S(0.5) / t;
      x() = (m.coeff(2, 1) - m.coeff(1, 2)) * t;
      y() = (m.coeff(0, 2) - m.coeff(2, 0)) * t;
      z() = (m.coeff(1, 0) - m.coeff(0, 1)) * t;
    } else {
      Index i = 0;
      if (m.coeff(1, 1) > m.coeff(0, 0)) i = 1;
      if (m.coeff(2, 2) > m.coeff(i, i)) i = 2;
      const Index j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m.coeff(i, i) - m.coeff(j, j) - m.coeff(k, k) + S(1.0));
      c_[i] = S(0.5) * t;
      t = S(0.5) / t;
      w() = (m.coeff(k, j) - m.coeff(j, k)) * t;
      c_[j] = (m.coeff(j, i) + m.coeff(i, j)) * t;
      c_[k] = (m.coeff(k, i) + m.coeff(i, k)) * t;
    }
  }
  Coefficients c_;
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

template <class S> class AngleAxis {
public:
  typedef Matrix<S, 3, 1> Vector3;
  typedef Matrix<S, 3, 3> Matrix3;
  AngleAxis() : angle_(0) {}
  template <class D> AngleAxis(const S& a, const MatrixBase<D>& ax) : angle_(a), axis_(ax) {}
  explicit AngleAxis(const Quaternion<S>& q) {
    S n = q.vec().norm();
    if (n < std::numeric_limits<S>::epsilon()) n = Vector3(q.x(), q.y(), q.z()).norm();
    if (n > S(0)) { angle_ = S(2) * std::atan2(n, std::abs(q.w())); if (q.w() < 0) n = -n; axis_ = Vector3(q.x(), q.y(), q.z()) / n; }
    else { angle_ = S(0); axis_ = Vector3(1, 0, 0); }
  }
  template <class D> explicit AngleAxis(const MatrixBase<D>& m) { *this = AngleAxis(Quaternion<S>(m)); }
  S angle() const { return angle_; }
  S& angle() { return angle_; }
  const Vector3& axis() const { return axis_; }
  Vector3& axis() { return axis_; }
  Matrix3 toRotationMatrix() const {
    Matrix3 r;
    const S s = std::sin(angle_), c = std::cos(angle_);
    Vector3 sin_axis = s * axis_;
    Vector3 cos1_axis = (S(1) - c) * axis_;
    S tmp;
    tmp = cos1_axis.x() * axis_.y(); r(0, 1) = tmp - sin_axis.z(); r(1, 0) = tmp + sin_axis.z();
    tmp = cos1_axis.x() * axis_.z(); r(0, 2) = tmp + sin_axis.y(); r(2, 0) = tmp - sin_axis.y();
    tmp = cos1_axis.y() * axis_.z(); r(1, 2) = tmp - sin_axis.x(); r(2, 1) = tmp + sin_axis.x();
    Vector3 d = cos1_axis.cwiseProduct(axis_);
    r(0, 0) = d[0] + c; r(1, 1) = d[1] + c; r(2, 2) = d[2] + c;
    return r;
  }
  Matrix3 matrix() const { return toRotationMatrix(); }
  template <class D> Vector3 operator*(const MatrixBase<D>& v) const { return toRotationMatrix() * v; }
private:
  S angle_;
  Vector3 axis_;
};
typedef AngleAxis<double> AngleAxisd;
typedef AngleAxis<float> AngleAxisf;
template <class S> Quaternion<S>::Quaternion(const AngleAxis<S>& aa) {
  const S ha = S(0.5) * aa.angle();
  w() = std::cos(ha);
  const S s = std::sin(ha);
  c_[0] = s * aa.axis()[0]; c_[1] = s * aa.axis()[1]; c_[2] = s * aa.axis()[2];
}

template <class S, int Dim, int Mode = Affine, int Opt = 0> class Transform {
public:
  typedef Matrix<S, Dim + 1, Dim + 1> MatrixType;
  typedef Matrix<S, Dim, Dim> LinearMatrixType;
  typedef Matrix<S, Dim, 1> VectorType;
  Transform() { m_.setIdentity(); }
  template <class D> explicit Transform(const MatrixBase<D>& m) { m_.setIdentity(); if (m.rows() == Dim) m_.template block<Dim, Dim>(0, 0) = m; else m_ = m; }
  Transform(const Quaternion<S>& q) { m_.setIdentity(); m_.template block<Dim, Dim>(0, 0) = q.toRotationMatrix(); }
  template <class D> Transform& operator=(const MatrixBase<D>& m) { m_.setIdentity(); if (m.rows() == Dim) m_.template block<Dim, Dim>(0, 0) = m; else m_ = m; return *this; }
  Transform& operator=(const Quaternion<S>& q) { m_.setIdentity(); m_.template block<Dim, Dim>(0, 0) = q.toRotationMatrix(); return *this; }
  static Transform Identity() { return Transform(); }
  void setIdentity() { m_.setIdentity(); }
  MatrixType& matrix() { return m_; }
  const MatrixType& matrix() const { return m_; }
  Block<S, Dim, Dim> linear() { return m_.template block<Dim, Dim>(0, 0); }
  Block<const S, Dim, Dim> linear() const { return m_.template block<Dim, Dim>(0, 0); }
  LinearMatrixType rotation() const { return LinearMatrixType(linear()); }
  Block<S, Dim, 1> translation() { return m_.template block<Dim, 1>(0, Dim); }
  Block<const S, Dim, 1> translation() const { return m_.template block<Dim, 1>(0, Dim); }
  S& operator()(Index i, Index j) { return m_(i, j); }
  const S& operator()(Index i, Index j) const { return m_(i, j); }
  S* data() { return m_.data(); }
  const S* data() const { return m_.data(); }
  Transform operator*(const Transform& o) const { Transform t; t.m_ = m_ * o.m_; return t; }
  Transform& operator*=(const Transform& o) { m_ = m_ * o.m_; return *this; }
  template <class D> VectorType operator*(const MatrixBase<D>& v) const { return VectorType(LinearMatrixType(linear()) * v + VectorType(translation())); }
  Transform inverse(int = Mode) const {
    Transform t;
    if (Mode == int(Isometry)) { LinearMatrixType rt = LinearMatrixType(linear()).transpose(); t.linear() = rt; t.translation() = -(rt * VectorType(translation())); }
    else { LinearMatrixType li = LinearMatrixType(linear()).inverse(); t.linear() = li; t.translation() = -(li * VectorType(translation())); }
    return t;
  }
  Transform& translate(const VectorType& v) { translation() += LinearMatrixType(linear()) * v; return *this; }
  Transform& pretranslate(const VectorType& v) { translation() += v; return *this; }
  template <class D> Transform& rotate(const MatrixBase<D>& r) { linear() = LinearMatrixType(linear()) * r; return *this; }
private:
  MatrixType m_;
};
typedef Transform<double, 3, Isometry> Isometry3d;
typedef Transform<double, 2, Isometry> Isometry2d;
typedef Transform<double, 3, Affine> Affine3d;
typedef Transform<double, 2, Affine> Affine2d;
typedef Transform<float, 3, Isometry> Isometry3f;
typedef Transform<float, 3, Affine> Affine3f;

}  // namespace Eigen
#endif
