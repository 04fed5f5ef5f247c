// eigenmini, sparse part -- TEST INFRASTRUCTURE (oracle/).  What Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h asks of
// <Eigen/Sparse> and <Eigen/SparseCholesky>: a compressed-column SparseMatrix built from triplets, a PermutationMatrix, and
// SimplicialLDLT<SparseMatrix, Upper> (analyzePattern / factorize / solve / info).  The factorisation is an unpivoted LDL^T of
// P A P^T like Eigen's; the fill-reducing ordering P is the identity unless the caller supplies one (Eigen's AMD ordering only
// changes the elimination order, i.e. the solution up to rounding).  The reduced camera systems this is used for are a few
// hundred unknowns, so the numeric phase works on a dense copy.
#ifndef ORBX_EIGENMINI_SPARSE_HPP
#define ORBX_EIGENMINI_SPARSE_HPP
#include "eigenmini.hpp"

namespace Eigen {

template <class S, class I = int> class Triplet {
public:
  Triplet() : r_(0), c_(0), v_(0) {}
  Triplet(const I& r, const I& c, const S& v = S(0)) : r_(r), c_(c), v_(v) {}
  const I& row() const { return r_; }
  const I& col() const { return c_; }
  const S& value() const { return v_; }
private:
  I r_, c_;
  S v_;
};

template <int SizeAtCompileTime = Dynamic, int MaxSize = SizeAtCompileTime, class I = int> class PermutationMatrix {
public:
  typedef Matrix<I, Dynamic, 1> IndicesType;
  PermutationMatrix() {}
  explicit PermutationMatrix(Index n) : idx_(n) {}
  void resize(Index n) { idx_.resize(n); }
  Index size() const { return idx_.size(); }
  Index rows() const { return idx_.size(); }
  Index cols() const { return idx_.size(); }
  IndicesType& indices() { return idx_; }
  const IndicesType& indices() const { return idx_; }
  void setIdentity() { for (Index i = 0; i < size(); ++i) idx_[i] = I(i); }
  void setIdentity(Index n) { resize(n); setIdentity(); }
  PermutationMatrix inverse() const { PermutationMatrix r(size()); for (Index i = 0; i < size(); ++i) r.idx_[idx_[i]] = I(i); return r; }
  PermutationMatrix transpose() const { return inverse(); }
private:
  IndicesType idx_;
};

template <class S, int Opt, class I> class SparseMatrix;
template <class SM, int UpLo> class SparseSelfAdjointView;
template <class SM, int UpLo> struct SparseSymmetricPermutation {
  const SM& m;
  const PermutationMatrix<Dynamic, Dynamic>& p;
};

template <class S, int Opt = ColMajor, class I = int> class SparseMatrix {
public:
  typedef S Scalar;
  typedef I StorageIndex;
  SparseMatrix() : r_(0), c_(0), outer_(1, 0) {}
  SparseMatrix(Index r, Index c) : r_(r), c_(c), outer_(size_t(c) + 1, 0) {}
  void resize(Index r, Index c) { r_ = r; c_ = c; outer_.assign(size_t(c) + 1, 0); inner_.clear(); val_.clear(); }
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Index nonZeros() const { return Index(val_.size()); }
  S* valuePtr() { return val_.data(); }
  const S* valuePtr() const { return val_.data(); }
  I* innerIndexPtr() { return inner_.data(); }
  const I* innerIndexPtr() const { return inner_.data(); }
  I* outerIndexPtr() { return outer_.data(); }
  const I* outerIndexPtr() const { return outer_.data(); }
  void setZero() { std::fill(outer_.begin(), outer_.end(), 0); inner_.clear(); val_.clear(); }
  // column-major compressed storage, row indices ascending inside a column, duplicates summed (Eigen's setFromTriplets)
  template <class It> void setFromTriplets(It b, It e) {
    std::vector<std::vector<std::pair<I, S> > > cols(static_cast<size_t>(c_));
    for (It t = b; t != e; ++t) cols[size_t(t->col())].push_back(std::make_pair(I(t->row()), S(t->value())));
    inner_.clear(); val_.clear();
    outer_.assign(size_t(c_) + 1, 0);
    for (Index c = 0; c < c_; ++c) {
      std::vector<std::pair<I, S> >& v = cols[size_t(c)];
      std::stable_sort(v.begin(), v.end(), [](const std::pair<I, S>& x, const std::pair<I, S>& y) { return x.first < y.first; });
      for (size_t k = 0; k < v.size(); ++k) {
        if (k && v[k].first == v[k - 1].first) val_.back() += v[k].second;
        else { inner_.push_back(v[k].first); val_.push_back(v[k].second); }
      }
      outer_[size_t(c) + 1] = I(val_.size());
    }
  }
  S coeff(Index r, Index c) const {
    for (I k = outer_[size_t(c)]; k < outer_[size_t(c) + 1]; ++k) if (inner_[size_t(k)] == r) return val_[size_t(k)];
    return S(0);
  }
  template <int UpLo> SparseSelfAdjointView<SparseMatrix, UpLo> selfadjointView() { return SparseSelfAdjointView<SparseMatrix, UpLo>(*this); }
  template <int UpLo> SparseSelfAdjointView<const SparseMatrix, UpLo> selfadjointView() const { return SparseSelfAdjointView<const SparseMatrix, UpLo>(*this); }
  // full symmetric matrix from a triangular view
  template <class SM, int UpLo> SparseMatrix& operator=(const SparseSelfAdjointView<SM, UpLo>& v) {
    const SM& a = v.matrix();
    std::vector<Triplet<S, I> > t;
    for (Index c = 0; c < a.cols(); ++c)
      for (I k = a.outerIndexPtr()[c]; k < a.outerIndexPtr()[c + 1]; ++k) {
        const I r = a.innerIndexPtr()[k];
        if ((UpLo == Upper && r > c) || (UpLo == Lower && r < c)) continue;
        t.push_back(Triplet<S, I>(r, I(c), a.valuePtr()[k]));
        if (r != c) t.push_back(Triplet<S, I>(I(c), r, a.valuePtr()[k]));
      }
    resize(a.rows(), a.cols());
    setFromTriplets(t.begin(), t.end());
    return *this;
  }
  const SparseMatrix& nestedExpression() const { return *this; }
private:
  Index r_, c_;
  std::vector<I> outer_, inner_;
  std::vector<S> val_;
};

template <class SM, int UpLo> class SparseSelfAdjointView {
public:
  typedef typename std::remove_const<SM>::type Plain;
  explicit SparseSelfAdjointView(SM& m) : m_(m) {}
  SM& matrix() const { return m_; }
  SparseSymmetricPermutation<Plain, UpLo> twistedBy(const PermutationMatrix<Dynamic, Dynamic>& p) const { return SparseSymmetricPermutation<Plain, UpLo>{m_, p}; }
  // dst.selfadjointView<DstUpLo>() = src.selfadjointView<SrcUpLo>().twistedBy(P):  dst = triangular part of P src P^-1
  template <int SrcUpLo> SparseSelfAdjointView& operator=(const SparseSymmetricPermutation<Plain, SrcUpLo>& sp) {
    typedef typename Plain::Scalar S;
    std::vector<Triplet<S, int> > t;
    const Plain& a = sp.m;
    for (Index c = 0; c < a.cols(); ++c)
      for (int k = a.outerIndexPtr()[c]; k < a.outerIndexPtr()[c + 1]; ++k) {
        const int r = a.innerIndexPtr()[k];
        if ((SrcUpLo == Upper && r > c) || (SrcUpLo == Lower && r < c)) continue;
        int pr = sp.p.indices()[r], pc = sp.p.indices()[c];
        if ((UpLo == Upper && pr > pc) || (UpLo == Lower && pr < pc)) std::swap(pr, pc);
        t.push_back(Triplet<S, int>(pr, pc, a.valuePtr()[k]));
      }
    m_.resize(a.rows(), a.cols());
    m_.setFromTriplets(t.begin(), t.end());
    return *this;
  }
private:
  SM& m_;
};

namespace internal {
// stand-in for Eigen's approximate-minimum-degree ordering: the identity (any ordering gives the same solution up to rounding)
template <class SM, class P> void minimum_degree_ordering(SM& c, P& perm) { perm.resize(c.cols()); for (Index i = 0; i < c.cols(); ++i) perm.indices()[i] = int(i); }
}  // namespace internal

template <class SM, int UpLo_ = Lower> class SimplicialLDLT {
public:
  typedef typename SM::Scalar Scalar;
  typedef SparseMatrix<Scalar, ColMajor, int> CholMatrixType;
  typedef Matrix<Scalar, Dynamic, 1> VectorType;
  enum { UpLo = UpLo_ };
  SimplicialLDLT() : info_(Success), n_(0) {}
  explicit SimplicialLDLT(const SM& a) : info_(Success), n_(0) { compute(a); }
  ComputationInfo info() const { return info_; }
  SimplicialLDLT& compute(const SM& a) { analyzePattern(a); factorize(a); return *this; }
  void analyzePattern(const SM& a) { n_ = a.cols(); m_P.resize(0); m_Pinv.resize(0); }
  void factorize(const SM& a) {
    n_ = a.cols();
    const Index n = n_;
    // dense copy of P A P^T (lower triangle)
    l_.resize(n, n);
    l_.setZero();
    const bool perm = m_Pinv.size() == n;
    for (Index c = 0; c < n; ++c)
      for (int k = a.outerIndexPtr()[c]; k < a.outerIndexPtr()[c + 1]; ++k) {
        const int r = a.innerIndexPtr()[k];
        if ((UpLo_ == Upper && r > c) || (UpLo_ == Lower && r < c)) continue;
        Index pr = perm ? m_Pinv.indices()[r] : r, pc = perm ? m_Pinv.indices()[c] : Index(c);
        if (pr < pc) std::swap(pr, pc);
        l_.coeffRef(pr, pc) = a.valuePtr()[k];
      }
    // up-looking LDL^T: row k of L from rows 0..k-1 (Eigen's factorize_preordered computes the same quantities column-sparse)
    d_.resize(n);
    info_ = Success;
    for (Index k = 0; k < n; ++k) {
      Scalar dk = l_.coeff(k, k);
      for (Index j = 0; j < k; ++j) {
        Scalar y = l_.coeff(k, j);
        for (Index i = 0; i < j; ++i) y -= l_.coeff(j, i) * l_.coeff(k, i) * d_[i];
        const Scalar lkj = y / d_[j];
        l_.coeffRef(k, j) = lkj;
      }
      for (Index j = 0; j < k; ++j) dk -= l_.coeff(k, j) * l_.coeff(k, j) * d_[j];
      d_[k] = dk;
      if (dk == Scalar(0)) { info_ = NumericalIssue; return; }
    }
  }
  template <class B> VectorType solve(const MatrixBase<B>& b) const {
    const Index n = n_;
    const bool perm = m_Pinv.size() == n;
    VectorType y(n);
    for (Index i = 0; i < n; ++i) y[perm ? Index(m_Pinv.indices()[i]) : i] = b.coeff(i);
    for (Index i = 0; i < n; ++i) { Scalar s = y[i]; for (Index k = 0; k < i; ++k) s -= l_.coeff(i, k) * y[k]; y[i] = s; }
    for (Index i = 0; i < n; ++i) y[i] /= d_[i];
    for (Index i = n - 1; i >= 0; --i) { Scalar s = y[i]; for (Index k = i + 1; k < n; ++k) s -= l_.coeff(k, i) * y[k]; y[i] = s; }
    VectorType x(n);
    for (Index i = 0; i < n; ++i) x[i] = y[perm ? Index(m_Pinv.indices()[i]) : i];
    return x;
  }
  struct LView {
    Index nnz;
    const LView& nestedExpression() const { return *this; }
    Index nonZeros() const { return nnz; }
  };
  LView matrixL() const { Index c = 0; for (Index j = 0; j < n_; ++j) for (Index i = j + 1; i < n_; ++i) if (l_.coeff(i, j) != Scalar(0)) ++c; return LView{c}; }
  VectorType vectorD() const { return d_; }
protected:
  void analyzePattern_preordered(const CholMatrixType& ap, bool) { n_ = ap.cols(); }
  PermutationMatrix<Dynamic, Dynamic> m_P, m_Pinv;
  ComputationInfo info_;
  Index n_;
  Matrix<Scalar, Dynamic, Dynamic> l_;
  VectorType d_;
};

template <class SM, int UpLo_ = Lower> class SimplicialLLT : public SimplicialLDLT<SM, UpLo_> {};

}  // namespace Eigen
#endif
