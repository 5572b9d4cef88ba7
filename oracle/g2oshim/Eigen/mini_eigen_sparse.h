/* oracle/g2oshim/Eigen/mini_eigen_sparse.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The part of Eigen's sparse interface that g2o's LinearSolverEigen (O3/Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h)
 * names: a compressed-column SparseMatrix filled from triplets (g2o later rewrites the values in place, in
 * compressed-column order), and SimplicialLDLT<.., Upper>.  The factorisation here expands the stored upper triangle to
 * a dense symmetric matrix and runs LDL^T without reordering: Eigen's AMD ordering only changes the rounding, and the
 * reduced camera systems of the bundle adjustments under test are a few hundred unknowns. */
#ifndef DVM_MINI_EIGEN_SPARSE_H
#define DVM_MINI_EIGEN_SPARSE_H
#include "mini_eigen.h"
#include <map>

namespace Eigen {

template <class T, int Options = ColMajor, class I = int> class SparseMatrix;

template <class SM, int UpLo_> class SparseSelfAdjointView {
public:
    SM& m;
    explicit SparseSelfAdjointView(SM& m_) : m(m_) { }
    template <class P> SparseSelfAdjointView twistedBy(const P&) const { return *this; }
    SparseSelfAdjointView(const SparseSelfAdjointView& o) : m(o.m) { }
    SparseSelfAdjointView& operator=(const SparseSelfAdjointView& o) { m = o.m; return *this; }
};

template <class T, int Options, class I> class SparseMatrix {
    int r_ = 0, c_ = 0;
    std::vector<I> outer_, inner_;
    std::vector<T> val_;
public:
    typedef T Scalar;
    typedef I StorageIndex;
    SparseMatrix() { }
    SparseMatrix(int r, int c) { resize(r, c); }
    void resize(int r, int c) { r_ = r; c_ = c; outer_.assign((size_t)c + 1, 0); inner_.clear(); val_.clear(); }
    int rows() const { return r_; }
    int cols() const { return c_; }
    int nonZeros() const { return (int)val_.size(); }
    T* valuePtr() { return val_.data(); }
    const T* valuePtr() const { return val_.data(); }
    I* innerIndexPtr() { return inner_.data(); }
    const I* innerIndexPtr() const { return inner_.data(); }
    I* outerIndexPtr() { return outer_.data(); }
    const I* outerIndexPtr() const { return outer_.data(); }
    void makeCompressed() { }
    template <class It> void setFromTriplets(It b, It e)
    {
        std::vector<std::map<I, T>> cols((size_t)c_);
        for (It it = b; it != e; ++it) cols[(size_t)it->col()][it->row()] += it->value();   /* duplicates are summed */
        outer_.assign((size_t)c_ + 1, 0); inner_.clear(); val_.clear();
        for (int c = 0; c < c_; c++) {
            for (const auto& kv : cols[(size_t)c]) { inner_.push_back(kv.first); val_.push_back(kv.second); }
            outer_[(size_t)c + 1] = (I)inner_.size();
        }
    }
    template <int U> SparseSelfAdjointView<SparseMatrix, U> selfadjointView() { return SparseSelfAdjointView<SparseMatrix, U>(*this); }
    template <int U> SparseSelfAdjointView<const SparseMatrix, U> selfadjointView() const { return SparseSelfAdjointView<const SparseMatrix, U>(*this); }
    template <class SM2, int U> SparseMatrix& operator=(const SparseSelfAdjointView<SM2, U>& v) { *this = v.m; return *this; }
    const SparseMatrix& nestedExpression() const { return *this; }
};

namespace internal {
template <class M, class P> void minimum_degree_ordering(M& m, P& perm)
{   /* identity: no fill-reducing ordering in this stand-in */
    perm.resize(m.cols());
    for (int i = 0; i < m.cols(); i++) perm.indices()(i) = i;
}
} // namespace internal

template <class SM, int UpLo_ = Lower> class SimplicialLDLT {
public:
    typedef SM MatrixType;
    typedef SM CholMatrixType;
    typedef typename SM::Scalar Scalar;
    enum { UpLo = UpLo_ };
    SimplicialLDLT() { }
    void analyzePattern(const SM&) { }
    void analyzePattern_preordered(const SM&, bool) { }
    void factorize(const SM& a)
    {
        const int n = a.cols();
        Matrix<Scalar, Dynamic, Dynamic> d(n, n);
        d.setZero();
        for (int c = 0; c < n; c++)
            for (int k = a.outerIndexPtr()[c]; k < a.outerIndexPtr()[c + 1]; k++) {
                const int r = a.innerIndexPtr()[k];
                d(r, c) = a.valuePtr()[k];
                d(c, r) = a.valuePtr()[k];
            }
        ldlt_.compute(d);
        /* SimplicialLDLT reports failure on a zero pivot only; g2o treats !Success as "not positive definite" */
        info_ = ldlt_.info();
        l_.resize(n, n);
    }
    void compute(const SM& a) { analyzePattern(a); factorize(a); }
    ComputationInfo info() const { return info_; }
    template <class D> Matrix<Scalar, Dynamic, 1> solve(const MatrixBase<D>& b) const
    {
        Matrix<Scalar, Dynamic, 1> bb;
        bb.assign(b.derived());
        return ldlt_.solve(bb);
    }
    const SM& matrixL() const { return l_; }
protected:
    PermutationMatrix<Dynamic, Dynamic, int> m_P, m_Pinv;
    LDLT<Matrix<Scalar, Dynamic, Dynamic>> ldlt_;
    ComputationInfo info_ = Success;
    SM l_;
};

} // namespace Eigen
#endif
