/* oracle/g2oshim/Eigen/mini_eigen.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A small dense linear-algebra library with the part of Eigen 3's interface that the reference's vendored g2o
 * (O3/Thirdparty/g2o) and the optimisation functions of O3/src/Optimizer.cc use, so that those sources can be
 * compiled UNMODIFIED where they lie (Eigen itself is not installed in this image and there is no network).
 * Value semantics throughout: every operation returns a plain Matrix, there are no expression templates, hence no
 * aliasing rules; Map<> and Block<> are views that can be assigned to.  Storage is column-major like Eigen's default
 * (g2o maps Hessian blocks over raw double arrays and relies on that).  Arithmetic is plain IEEE double evaluated in
 * the natural loop order -- g2o's results are compared within a stated tolerance, not bit for bit (the reference's own
 * summation order depends on allocation addresses).
 * Everything lives in namespace Eigen because the reference's sources name it. */
#ifndef DVM_MINI_EIGEN_H
#define DVM_MINI_EIGEN_H
#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstring>
#include <iostream>
#include <memory>
#include <type_traits>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW_IF(x)
#define EIGEN_DEFINE_STL_VECTOR_SPECIALIZATION(...)
#define EIGEN_WORLD_VERSION 3
#define EIGEN_MAJOR_VERSION 4
#define EIGEN_MINOR_VERSION 0
#define EIGEN_VERSION_AT_LEAST(x, y, z) 1
#define EIGEN_STRONG_INLINE inline

namespace Eigen {

typedef std::ptrdiff_t Index;
const int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 1, AutoAlign = 0, DontAlign = 2 };
enum { Lower = 1, Upper = 2 };
enum { Unaligned = 0, Aligned = 16 };
enum { AlignedBit = 0x40 };
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };
enum TransformTraits { Isometry = 1, Affine = 2, AffineCompact = 0x12, Projective = 0x20 };
enum { EigenvaluesOnly = 0x40, ComputeEigenvectors = 0x80 };
inline void initParallel() { }
template <class T> using aligned_allocator = std::allocator<T>;

template <class T, int R, int C, int O = 0, int MR = R, int MC = C> class Matrix;
template <class Derived> struct traits;
template <class XprType, int BR, int BC> class Block;
template <class PlainType, int MapOptions = 0> class Map;
template <class Derived> class ArrayView;
template <class XprType> class DiagonalView;
template <class MatrixType> class LDLT;
template <class MatrixType> class LLT;
template <class MatrixType> class PartialPivLU;

namespace internal {
template <class T, int R, int C> struct storage {   /* fixed size */
    T m[R * C];
    storage() { }   /* uninitialised like Eigen */
    int rows() const { return R; }
    int cols() const { return C; }
    void resize(int r, int c) { assert(r == R && c == C); (void)r; (void)c; }
    T* data() { return m; }
    const T* data() const { return m; }
};
template <class T, int R, int C> struct dyn_storage {
    std::vector<T> v;
    int r = (R == Dynamic ? 0 : R), c = (C == Dynamic ? 0 : C);
    int rows() const { return r; }
    int cols() const { return c; }
    void resize(int r_, int c_) { r = r_; c = c_; v.resize((size_t)r_ * c_); }
    T* data() { return v.data(); }
    const T* data() const { return v.data(); }
};
template <class T, int R, int C> struct pick_storage {
    typedef typename std::conditional<R == Dynamic || C == Dynamic, dyn_storage<T, R, C>, storage<T, R, C>>::type type;
};
template <int A, int B> struct prod_dim { enum { value = A }; };
template <class Derived> struct plain_of {
    typedef Matrix<typename traits<Derived>::Scalar, traits<Derived>::Rows, traits<Derived>::Cols> type;
};
template <class Derived> struct transposed_of {
    typedef Matrix<typename traits<Derived>::Scalar, traits<Derived>::Cols, traits<Derived>::Rows> type;
};
template <int A, int B> struct same_dim { enum { value = (A == Dynamic ? B : A) }; };
} // namespace internal

template <class T, int R, int C, int O, int MR, int MC> struct traits<Matrix<T, R, C, O, MR, MC>> {
    typedef T Scalar; enum { Rows = R, Cols = C };
};
template <class X, int BR, int BC> struct traits<Block<X, BR, BC>> {
    typedef typename traits<X>::Scalar Scalar; enum { Rows = BR, Cols = BC };
};
template <class P, int MO> struct traits<Map<P, MO>> {
    typedef typename traits<typename std::remove_const<P>::type>::Scalar Scalar;
    enum { Rows = traits<typename std::remove_const<P>::type>::Rows, Cols = traits<typename std::remove_const<P>::type>::Cols };
};

/* comma initialiser: m << a, b, c; row-major fill, accepts scalars and matrix blocks (vectors stacked / rows appended) */
template <class Derived> class CommaInitializer {
public:
    typedef typename traits<Derived>::Scalar Scalar;
    CommaInitializer(Derived& m, Scalar s) : m_(m), row_(0), col_(0), blockRows_(1) { put(s); }
    template <class O> CommaInitializer(Derived& m, const O& o, int) : m_(m), row_(0), col_(0), blockRows_(1) { putBlock(o); }
    CommaInitializer& operator,(Scalar s) { put(s); return *this; }
    template <class O> typename std::enable_if<!std::is_arithmetic<O>::value, CommaInitializer&>::type operator,(const O& o)
    { putBlock(o); return *this; }
private:
    void wrap() { if (col_ >= m_.cols()) { col_ = 0; row_ += blockRows_; blockRows_ = 1; } }
    void put(Scalar s) { wrap(); m_.coeffRef(row_, col_) = s; col_++; }
    template <class O> void putBlock(const O& o)
    {
        wrap();
        for (int i = 0; i < o.rows(); i++)
            for (int j = 0; j < o.cols(); j++) m_.coeffRef(row_ + i, col_ + j) = o.coeff(i, j);
        col_ += o.cols();
        blockRows_ = o.rows();
    }
    Derived& m_;
    int row_, col_, blockRows_;
};

template <class Derived> class MatrixBase {
public:
    typedef typename traits<Derived>::Scalar Scalar;
    typedef Scalar RealScalar;
    enum { RowsAtCompileTime = traits<Derived>::Rows, ColsAtCompileTime = traits<Derived>::Cols,
           SizeAtCompileTime = (traits<Derived>::Rows == Dynamic || traits<Derived>::Cols == Dynamic) ? Dynamic
                                                                                                      : traits<Derived>::Rows * traits<Derived>::Cols,
           IsVectorAtCompileTime = (traits<Derived>::Rows == 1 || traits<Derived>::Cols == 1) };
    typedef typename internal::plain_of<Derived>::type PlainObject;
    typedef typename internal::transposed_of<Derived>::type TransposedObject;
    Derived& derived() { return *static_cast<Derived*>(this); }
    const Derived& derived() const { return *static_cast<const Derived*>(this); }
    int rows() const { return derived().rows(); }
    int cols() const { return derived().cols(); }
    int size() const { return rows() * cols(); }
    Scalar coeff(int i, int j) const { return derived().coeff(i, j); }
    Scalar& coeffRef(int i, int j) { return derived().coeffRef(i, j); }
    Scalar operator()(int i, int j) const { return coeff(i, j); }
    Scalar& operator()(int i, int j) { return coeffRef(i, j); }
    /* vector access (column or row vectors) */
    Scalar coeff(int i) const { return cols() == 1 ? coeff(i, 0) : coeff(0, i); }
    Scalar& coeffRef(int i) { return cols() == 1 ? coeffRef(i, 0) : coeffRef(0, i); }
    Scalar operator()(int i) const { return coeff(i); }
    Scalar& operator()(int i) { return coeffRef(i); }
    Scalar operator[](int i) const { return coeff(i); }
    Scalar& operator[](int i) { return coeffRef(i); }
    Scalar x() const { return coeff(0); } Scalar y() const { return coeff(1); } Scalar z() const { return coeff(2); } Scalar w() const { return coeff(3); }
    Scalar& x() { return coeffRef(0); } Scalar& y() { return coeffRef(1); } Scalar& z() { return coeffRef(2); } Scalar& w() { return coeffRef(3); }

    PlainObject eval() const { PlainObject p; p.assign(derived()); return p; }
    Derived& noalias() { return derived(); }
    template <class U> Matrix<U, traits<Derived>::Rows, traits<Derived>::Cols> cast() const
    {
        Matrix<U, traits<Derived>::Rows, traits<Derived>::Cols> o;
        o.resize(rows(), cols());
        for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) o.coeffRef(i, j) = (U)coeff(i, j);
        return o;
    }
    TransposedObject transpose() const
    {
        TransposedObject o;
        o.resize(cols(), rows());
        for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) o.coeffRef(j, i) = coeff(i, j);
        return o;
    }
    TransposedObject adjoint() const { return transpose(); }
    void transposeInPlace() { PlainObject t; t.assign(transpose()); derived() = t; }

    /* fills */
    Derived& setZero() { return setConstant(Scalar(0)); }
    Derived& setOnes() { return setConstant(Scalar(1)); }
    Derived& setConstant(Scalar s) { for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) coeffRef(i, j) = s; return derived(); }
    void fill(Scalar s) { setConstant(s); }
    Derived& setIdentity() { for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) coeffRef(i, j) = Scalar(i == j); return derived(); }

    /* reductions */
    Scalar squaredNorm() const { Scalar s = 0; for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) s += coeff(i, j) * coeff(i, j); return s; }
    Scalar norm() const { return std::sqrt(squaredNorm()); }
    Scalar sum() const { Scalar s = 0; for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) s += coeff(i, j); return s; }
    Scalar trace() const { Scalar s = 0; for (int i = 0; i < std::min(rows(), cols()); i++) s += coeff(i, i); return s; }
    Scalar maxCoeff() const { Scalar s = coeff(0, 0); for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) s = std::max(s, coeff(i, j)); return s; }
    Scalar minCoeff() const { Scalar s = coeff(0, 0); for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) s = std::min(s, coeff(i, j)); return s; }
    bool hasNaN() const { for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) if (std::isnan(coeff(i, j))) return true; return false; }
    bool allFinite() const { for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) if (!std::isfinite(coeff(i, j))) return false; return true; }
    template <class O> Scalar dot(const MatrixBase<O>& o) const { Scalar s = 0; for (int i = 0; i < size(); i++) s += coeff(i) * o.coeff(i); return s; }
    template <class O> Matrix<Scalar, 3, 1> cross(const MatrixBase<O>& o) const
    {
        Matrix<Scalar, 3, 1> r;
        r.coeffRef(0, 0) = coeff(1) * o.coeff(2) - coeff(2) * o.coeff(1);
        r.coeffRef(1, 0) = coeff(2) * o.coeff(0) - coeff(0) * o.coeff(2);
        r.coeffRef(2, 0) = coeff(0) * o.coeff(1) - coeff(1) * o.coeff(0);
        return r;
    }
    void normalize() { const Scalar n = norm(); if (n > Scalar(0)) derived() /= n; }
    PlainObject normalized() const { PlainObject p = eval(); p.normalize(); return p; }
    PlainObject cwiseAbs() const { PlainObject p = eval(); for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) p.coeffRef(i, j) = std::abs(coeff(i, j)); return p; }
    template <class O> PlainObject cwiseProduct(const MatrixBase<O>& o) const
    { PlainObject p = eval(); for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) p.coeffRef(i, j) *= o.coeff(i, j); return p; }
    template <class O> bool isApprox(const MatrixBase<O>& o, Scalar prec = Scalar(1e-12)) const
    { return (eval() - o.eval()).squaredNorm() <= prec * prec * std::min(squaredNorm(), o.squaredNorm()); }

    /* views */
    template <int BR, int BC> Block<Derived, BR, BC> block(int i, int j) { return Block<Derived, BR, BC>(derived(), i, j, BR, BC); }
    template <int BR, int BC> Block<const Derived, BR, BC> block(int i, int j) const { return Block<const Derived, BR, BC>(derived(), i, j, BR, BC); }
    Block<Derived, Dynamic, Dynamic> block(int i, int j, int r, int c) { return Block<Derived, Dynamic, Dynamic>(derived(), i, j, r, c); }
    Block<const Derived, Dynamic, Dynamic> block(int i, int j, int r, int c) const { return Block<const Derived, Dynamic, Dynamic>(derived(), i, j, r, c); }
    template <int BR, int BC> Block<Derived, BR, BC> topLeftCorner() { return block<BR, BC>(0, 0); }
    template <int BR, int BC> Block<const Derived, BR, BC> topLeftCorner() const { return block<BR, BC>(0, 0); }
    template <int BR, int BC> Block<Derived, BR, BC> topRightCorner() { return block<BR, BC>(0, cols() - BC); }
    template <int BR, int BC> Block<const Derived, BR, BC> topRightCorner() const { return block<BR, BC>(0, cols() - BC); }
    template <int BR, int BC> Block<Derived, BR, BC> bottomLeftCorner() { return block<BR, BC>(rows() - BR, 0); }
    template <int BR, int BC> Block<Derived, BR, BC> bottomRightCorner() { return block<BR, BC>(rows() - BR, cols() - BC); }
    Block<Derived, Dynamic, Dynamic> topLeftCorner(int r, int c) { return block(0, 0, r, c); }
    Block<Derived, traits<Derived>::Rows, 1> col(int j) { return Block<Derived, traits<Derived>::Rows, 1>(derived(), 0, j, rows(), 1); }
    Block<const Derived, traits<Derived>::Rows, 1> col(int j) const { return Block<const Derived, traits<Derived>::Rows, 1>(derived(), 0, j, rows(), 1); }
    Block<Derived, 1, traits<Derived>::Cols> row(int i) { return Block<Derived, 1, traits<Derived>::Cols>(derived(), i, 0, 1, cols()); }
    Block<const Derived, 1, traits<Derived>::Cols> row(int i) const { return Block<const Derived, 1, traits<Derived>::Cols>(derived(), i, 0, 1, cols()); }
    /* vector segments (column vectors; row vectors handled by the orientation test) */
    template <int N> Block<Derived, (traits<Derived>::Cols == 1 ? N : 1), (traits<Derived>::Cols == 1 ? 1 : N)> segment(int i)
    { return seg_<N>(i, N); }
    template <int N> Block<const Derived, (traits<Derived>::Cols == 1 ? N : 1), (traits<Derived>::Cols == 1 ? 1 : N)> segment(int i) const
    { return cseg_<N>(i, N); }
    template <int N> Block<Derived, (traits<Derived>::Cols == 1 ? N : 1), (traits<Derived>::Cols == 1 ? 1 : N)> segment(int i, int n)
    { return seg_<N>(i, n); }
    template <int N> Block<const Derived, (traits<Derived>::Cols == 1 ? N : 1), (traits<Derived>::Cols == 1 ? 1 : N)> segment(int i, int n) const
    { return cseg_<N>(i, n); }
    Block<Derived, (traits<Derived>::Cols == 1 ? Dynamic : 1), (traits<Derived>::Cols == 1 ? 1 : Dynamic)> segment(int i, int n) { return seg_<Dynamic>(i, n); }
    Block<const Derived, (traits<Derived>::Cols == 1 ? Dynamic : 1), (traits<Derived>::Cols == 1 ? 1 : Dynamic)> segment(int i, int n) const { return cseg_<Dynamic>(i, n); }
    template <int N> auto head() -> decltype(this->template segment<N>(0)) { return segment<N>(0); }
    template <int N> auto head() const -> decltype(this->template segment<N>(0)) { return segment<N>(0); }
    template <int N> auto tail() -> decltype(this->template segment<N>(0)) { return segment<N>(size() - N); }
    template <int N> auto tail() const -> decltype(this->template segment<N>(0)) { return segment<N>(size() - N); }
    auto head(int n) -> decltype(this->segment(0, 0)) { return segment(0, n); }
    auto head(int n) const -> decltype(this->segment(0, 0)) { return segment(0, n); }
    auto tail(int n) -> decltype(this->segment(0, 0)) { return segment(size() - n, n); }
    auto tail(int n) const -> decltype(this->segment(0, 0)) { return segment(size() - n, n); }
    DiagonalView<Derived> diagonal() { return DiagonalView<Derived>(derived()); }
    DiagonalView<const Derived> diagonal() const { return DiagonalView<const Derived>(derived()); }
    ArrayView<Derived> array() { return ArrayView<Derived>(derived()); }
    Derived& matrix() { return derived(); }
    const Derived& matrix() const { return derived(); }

    /* compound assignment */
    template <class O> Derived& operator+=(const MatrixBase<O>& o) { for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) coeffRef(i, j) += o.coeff(i, j); return derived(); }
    template <class O> Derived& operator-=(const MatrixBase<O>& o) { for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) coeffRef(i, j) -= o.coeff(i, j); return derived(); }
    Derived& operator*=(Scalar s) { for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) coeffRef(i, j) *= s; return derived(); }
    Derived& operator/=(Scalar s) { for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) coeffRef(i, j) /= s; return derived(); }
    template <class O> Derived& operator*=(const MatrixBase<O>& o) { PlainObject t = (*this) * o; derived() = t; return derived(); }
    CommaInitializer<Derived> operator<<(Scalar s) { return CommaInitializer<Derived>(derived(), s); }
    template <class O> CommaInitializer<Derived> operator<<(const MatrixBase<O>& o) { return CommaInitializer<Derived>(derived(), o.derived(), 0); }

    /* dense decompositions and small inverses */
    PlainObject inverse() const;
    Scalar determinant() const;
    LDLT<PlainObject> ldlt() const;
    LLT<PlainObject> llt() const;
    PartialPivLU<PlainObject> lu() const;
    PartialPivLU<PlainObject> partialPivLu() const;

protected:
    template <int N> Block<Derived, (traits<Derived>::Cols == 1 ? N : 1), (traits<Derived>::Cols == 1 ? 1 : N)> seg_(int i, int n)
    {
        typedef Block<Derived, (traits<Derived>::Cols == 1 ? N : 1), (traits<Derived>::Cols == 1 ? 1 : N)> B;
        return traits<Derived>::Cols == 1 ? B(derived(), i, 0, n, 1) : B(derived(), 0, i, 1, n);
    }
    template <int N> Block<const Derived, (traits<Derived>::Cols == 1 ? N : 1), (traits<Derived>::Cols == 1 ? 1 : N)> cseg_(int i, int n) const
    {
        typedef Block<const Derived, (traits<Derived>::Cols == 1 ? N : 1), (traits<Derived>::Cols == 1 ? 1 : N)> B;
        return traits<Derived>::Cols == 1 ? B(derived(), i, 0, n, 1) : B(derived(), 0, i, 1, n);
    }
};

template <class X, int BR, int BC> struct traits<Block<const X, BR, BC>> {
    typedef typename traits<X>::Scalar Scalar; enum { Rows = BR, Cols = BC };
};
template <class X> struct traits<DiagonalView<X>> {
    typedef typename traits<typename std::remove_const<X>::type>::Scalar Scalar;
    enum { Rows = traits<typename std::remove_const<X>::type>::Rows, Cols = 1 };
};

/* ---- plain matrix ---- */
template <class T, int R, int C, int O, int MR, int MC> class Matrix : public MatrixBase<Matrix<T, R, C, O, MR, MC>> {
    typename internal::pick_storage<T, R, C>::type s_;
public:
    typedef MatrixBase<Matrix> Base;
    typedef T Scalar;
    enum { Flags = AlignedBit, Options = O };
    typedef Map<Matrix, 0> MapType;
    typedef Map<const Matrix, 0> ConstMapType;
    typedef Map<Matrix, Aligned> AlignedMapType;
    typedef Map<const Matrix, Aligned> ConstAlignedMapType;
    using Base::operator=; /* (none in base; kept for symmetry) */
    Matrix() { }
    explicit Matrix(int n) { if (R == Dynamic && C == 1) s_.resize(n, 1); else if (C == Dynamic && R == 1) s_.resize(1, n); else if (R == Dynamic && C == Dynamic) s_.resize(n, n); else if (R * C == 1) s_.data()[0] = (T)n; }
    /* two arguments: (rows, cols) of a dynamic matrix, the two coefficients of a fixed 2-vector */
    template <class A, class B, class = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>::type>
    Matrix(A a, B b) { init2(a, b, std::integral_constant<bool, (R == Dynamic || C == Dynamic)>()); }
    Matrix(T a, T b, T c) { s_.data()[0] = a; s_.data()[1] = b; s_.data()[2] = c; }
    Matrix(T a, T b, T c, T d) { s_.data()[0] = a; s_.data()[1] = b; s_.data()[2] = c; s_.data()[3] = d; }
    explicit Matrix(const T* p) { std::memcpy(s_.data(), p, sizeof(T) * R * C); }
    Matrix(const Matrix& o) : s_(o.s_) { }
    template <class D> Matrix(const MatrixBase<D>& o) { assign(o.derived()); }
    Matrix& operator=(const Matrix& o) { s_ = o.s_; return *this; }
    template <class D> Matrix& operator=(const MatrixBase<D>& o) { Matrix t; t.assign(o.derived()); s_ = t.s_; return *this; }
    template <class D> void assign(const D& o)
    {
        s_.resize(o.rows(), o.cols());
        for (int j = 0; j < o.cols(); j++) for (int i = 0; i < o.rows(); i++) coeffRef(i, j) = o.coeff(i, j);
    }
    int rows() const { return s_.rows(); }
    int cols() const { return s_.cols(); }
    T coeff(int i, int j) const { return s_.data()[(size_t)j * s_.rows() + i]; }
    T& coeffRef(int i, int j) { return s_.data()[(size_t)j * s_.rows() + i]; }
    using Base::coeff; using Base::coeffRef;
    T* data() { return s_.data(); }
    const T* data() const { return s_.data(); }
    void resize(int r, int c) { s_.resize(r, c); }
    void resize(int n) { if (C == 1) s_.resize(n, 1); else if (R == 1) s_.resize(1, n); else s_.resize(n, n); }
    void conservativeResize(int r, int c) { Matrix t(*this); s_.resize(r, c); Base::setZero(); for (int j = 0; j < std::min(c, t.cols()); j++) for (int i = 0; i < std::min(r, t.rows()); i++) coeffRef(i, j) = t.coeff(i, j); }
    void conservativeResize(int n) { if (C == 1) conservativeResize(n, 1); else conservativeResize(1, n); }
    void swap(Matrix& o) { std::swap(s_, o.s_); }
    static Matrix Zero() { Matrix m; m.setZero(); return m; }
    static Matrix Zero(int n) { Matrix m(n); m.setZero(); return m; }
    static Matrix Zero(int r, int c) { Matrix m(r, c); m.setZero(); return m; }
    static Matrix Ones() { Matrix m; m.setOnes(); return m; }
    static Matrix Ones(int n) { Matrix m(n); m.setOnes(); return m; }
    static Matrix Constant(T v) { Matrix m; m.setConstant(v); return m; }
    static Matrix Constant(int n, T v) { Matrix m(n); m.setConstant(v); return m; }
    static Matrix Constant(int r, int c, T v) { Matrix m(r, c); m.setConstant(v); return m; }
    static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
    static Matrix Identity(int r, int c) { Matrix m(r, c); m.setIdentity(); return m; }
    static Matrix Random() { Matrix m; for (int i = 0; i < m.size(); i++) m.data()[i] = (T)(2.0 * std::rand() / RAND_MAX - 1.0); return m; }
    static Matrix UnitX() { Matrix m; m.setZero(); m.coeffRef(0) = 1; return m; }
    static Matrix UnitY() { Matrix m; m.setZero(); m.coeffRef(1) = 1; return m; }
    static Matrix UnitZ() { Matrix m; m.setZero(); m.coeffRef(2) = 1; return m; }
private:
    template <class A, class B> void init2(A a, B b, std::false_type) { s_.data()[0] = (T)a; s_.data()[1] = (T)b; }
    template <class A, class B> void init2(A a, B b, std::true_type) { s_.resize((int)a, (int)b); }
};

/* ---- Map: a view over caller memory ---- */
template <class P, int MO> class Map : public MatrixBase<Map<P, MO>> {
    typedef typename std::remove_const<P>::type Plain;
    typedef typename traits<Plain>::Scalar T;
    typedef typename std::conditional<std::is_const<P>::value, const T*, T*>::type Ptr;
    Ptr p_;
    int r_, c_;
public:
    typedef MatrixBase<Map> Base;
    Map(Ptr p) : p_(p), r_(traits<Plain>::Rows), c_(traits<Plain>::Cols) { }
    Map(Ptr p, int n) : p_(p), r_(traits<Plain>::Cols == 1 ? n : (traits<Plain>::Rows == Dynamic ? n : traits<Plain>::Rows)),
                        c_(traits<Plain>::Cols == 1 ? 1 : n) { if (traits<Plain>::Rows == 1) { r_ = 1; c_ = n; } }
    Map(Ptr p, int r, int c) : p_(p), r_(r), c_(c) { }
    Map(const Map& o) : p_(o.p_), r_(o.r_), c_(o.c_) { }
    int rows() const { return r_; }
    int cols() const { return c_; }
    T coeff(int i, int j) const { return p_[(size_t)j * r_ + i]; }
    T& coeffRef(int i, int j) { return const_cast<T*>(p_)[(size_t)j * r_ + i]; }
    using Base::coeff; using Base::coeffRef;
    Ptr data() const { return p_; }
    Map& operator=(const Map& o) { return assignFrom(o); }
    template <class D> Map& operator=(const MatrixBase<D>& o) { Plain t; t.assign(o.derived()); return assignFrom(t); }
private:
    template <class D> Map& assignFrom(const D& o) { for (int j = 0; j < c_; j++) for (int i = 0; i < r_; i++) coeffRef(i, j) = o.coeff(i, j); return *this; }
};

/* ---- Block: a view into another expression ---- */
template <class X, int BR, int BC> class Block : public MatrixBase<Block<X, BR, BC>> {
    typedef typename std::remove_const<X>::type XP;
    typedef typename traits<XP>::Scalar T;
    X& x_;
    int i0_, j0_, r_, c_;
public:
    typedef MatrixBase<Block> Base;
    Block(X& x, int i0, int j0, int r, int c) : x_(x), i0_(i0), j0_(j0), r_(r), c_(c) { }
    Block(const Block& o) : x_(o.x_), i0_(o.i0_), j0_(o.j0_), r_(o.r_), c_(o.c_) { }
    int rows() const { return r_; }
    int cols() const { return c_; }
    T coeff(int i, int j) const { return x_.coeff(i0_ + i, j0_ + j); }
    T& coeffRef(int i, int j) { return const_cast<XP&>(x_).coeffRef(i0_ + i, j0_ + j); }
    using Base::coeff; using Base::coeffRef;
    Block& operator=(const Block& o) { typename Base::PlainObject t; t.assign(o); return assignFrom(t); }
    template <class D> Block& operator=(const MatrixBase<D>& o) { typename internal::plain_of<D>::type t; t.assign(o.derived()); return assignFrom(t); }
private:
    template <class D> Block& assignFrom(const D& o) { for (int j = 0; j < c_; j++) for (int i = 0; i < r_; i++) coeffRef(i, j) = o.coeff(i, j); return *this; }
};

template <class X> class DiagonalView : public MatrixBase<DiagonalView<X>> {
    typedef typename std::remove_const<X>::type XP;
    typedef typename traits<XP>::Scalar T;
    X& x_;
public:
    typedef MatrixBase<DiagonalView> Base;
    explicit DiagonalView(X& x) : x_(x) { }
    int rows() const { return std::min(x_.rows(), x_.cols()); }
    int cols() const { return 1; }
    T coeff(int i, int) const { return x_.coeff(i, i); }
    T& coeffRef(int i, int) { return const_cast<XP&>(x_).coeffRef(i, i); }
    using Base::coeff; using Base::coeffRef;
    DiagonalView& operator=(const DiagonalView& o) { for (int i = 0; i < rows(); i++) coeffRef(i, 0) = o.coeff(i, 0); return *this; }
    template <class D> DiagonalView& operator=(const MatrixBase<D>& o) { for (int i = 0; i < rows(); i++) coeffRef(i, 0) = o.coeff(i); return *this; }
};
template <class X> class ArrayView {
    X& x_;
public:
    explicit ArrayView(X& x) : x_(x) { }
    ArrayView& operator+=(typename traits<X>::Scalar s) { for (int j = 0; j < x_.cols(); j++) for (int i = 0; i < x_.rows(); i++) x_.coeffRef(i, j) += s; return *this; }
    ArrayView& operator-=(typename traits<X>::Scalar s) { return (*this) += -s; }
    ArrayView& operator*=(typename traits<X>::Scalar s) { for (int j = 0; j < x_.cols(); j++) for (int i = 0; i < x_.rows(); i++) x_.coeffRef(i, j) *= s; return *this; }
};

/* ---- arithmetic (results are plain matrices) ---- */
template <class A, class B>
Matrix<typename traits<A>::Scalar, internal::same_dim<traits<A>::Rows, traits<B>::Rows>::value, internal::same_dim<traits<A>::Cols, traits<B>::Cols>::value>
operator+(const MatrixBase<A>& a, const MatrixBase<B>& b)
{
    Matrix<typename traits<A>::Scalar, internal::same_dim<traits<A>::Rows, traits<B>::Rows>::value, internal::same_dim<traits<A>::Cols, traits<B>::Cols>::value> o;
    o.resize(a.rows(), a.cols());
    for (int j = 0; j < a.cols(); j++) for (int i = 0; i < a.rows(); i++) o.coeffRef(i, j) = a.coeff(i, j) + b.coeff(i, j);
    return o;
}
template <class A, class B>
Matrix<typename traits<A>::Scalar, internal::same_dim<traits<A>::Rows, traits<B>::Rows>::value, internal::same_dim<traits<A>::Cols, traits<B>::Cols>::value>
operator-(const MatrixBase<A>& a, const MatrixBase<B>& b)
{
    Matrix<typename traits<A>::Scalar, internal::same_dim<traits<A>::Rows, traits<B>::Rows>::value, internal::same_dim<traits<A>::Cols, traits<B>::Cols>::value> o;
    o.resize(a.rows(), a.cols());
    for (int j = 0; j < a.cols(); j++) for (int i = 0; i < a.rows(); i++) o.coeffRef(i, j) = a.coeff(i, j) - b.coeff(i, j);
    return o;
}
template <class A> typename internal::plain_of<A>::type operator-(const MatrixBase<A>& a)
{
    typename internal::plain_of<A>::type o;
    o.resize(a.rows(), a.cols());
    for (int j = 0; j < a.cols(); j++) for (int i = 0; i < a.rows(); i++) o.coeffRef(i, j) = -a.coeff(i, j);
    return o;
}
template <class A, class B>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> operator*(const MatrixBase<A>& a, const MatrixBase<B>& b)
{
    Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> o;
    o.resize(a.rows(), b.cols());
    const int K = a.cols();
    for (int j = 0; j < b.cols(); j++)
        for (int i = 0; i < a.rows(); i++) {
            typename traits<A>::Scalar s = 0;
            for (int k = 0; k < K; k++) s += a.coeff(i, k) * b.coeff(k, j);
            o.coeffRef(i, j) = s;
        }
    return o;
}
#define DVM_MINI_EIGEN_SCALAR_OPS(ST)                                                                                         \
    template <class A> typename std::enable_if<std::is_same<typename traits<A>::Scalar, ST>::value || true, typename internal::plain_of<A>::type>::type \
    operator*(const MatrixBase<A>& a, ST s)                                                                                   \
    { typename internal::plain_of<A>::type o = a.eval(); o *= (typename traits<A>::Scalar)s; return o; }                      \
    template <class A> typename internal::plain_of<A>::type operator*(ST s, const MatrixBase<A>& a)                           \
    { typename internal::plain_of<A>::type o = a.eval(); o *= (typename traits<A>::Scalar)s; return o; }                      \
    template <class A> typename internal::plain_of<A>::type operator/(const MatrixBase<A>& a, ST s)                           \
    { typename internal::plain_of<A>::type o = a.eval(); o /= (typename traits<A>::Scalar)s; return o; }
DVM_MINI_EIGEN_SCALAR_OPS(double)
DVM_MINI_EIGEN_SCALAR_OPS(float)
DVM_MINI_EIGEN_SCALAR_OPS(int)
#undef DVM_MINI_EIGEN_SCALAR_OPS
template <class A, class B> bool operator==(const MatrixBase<A>& a, const MatrixBase<B>& b)
{
    if (a.rows() != b.rows() || a.cols() != b.cols()) return false;
    for (int j = 0; j < a.cols(); j++) for (int i = 0; i < a.rows(); i++) if (a.coeff(i, j) != b.coeff(i, j)) return false;
    return true;
}
template <class A, class B> bool operator!=(const MatrixBase<A>& a, const MatrixBase<B>& b) { return !(a == b); }
template <class A> std::ostream& operator<<(std::ostream& os, const MatrixBase<A>& a)
{
    for (int i = 0; i < a.rows(); i++) {
        for (int j = 0; j < a.cols(); j++) os << (j ? " " : "") << a.coeff(i, j);
        if (i + 1 < a.rows()) os << "\n";
    }
    return os;
}

/* ---- decompositions ---- */
/* LDL^T without pivoting (Eigen's LDLT pivots; for the positive definite, damped systems of g2o both give the same
 * solution up to rounding) */
template <class M> class LDLT {
    typedef typename traits<M>::Scalar T;
    M a_;
    bool ok_ = false;
public:
    LDLT() { }
    template <class D> explicit LDLT(const MatrixBase<D>& a) { compute(a); }
    template <class D> LDLT& compute(const MatrixBase<D>& a)
    {
        a_.assign(a.derived());
        const int n = a_.rows();
        ok_ = true;
        for (int j = 0; j < n; j++) {
            T d = a_.coeff(j, j);
            for (int k = 0; k < j; k++) d -= a_.coeff(j, k) * a_.coeff(j, k) * a_.coeff(k, k);
            a_.coeffRef(j, j) = d;
            if (!(d > T(0)) && !(d < T(0))) ok_ = false;
            for (int i = j + 1; i < n; i++) {
                T s = a_.coeff(i, j);
                for (int k = 0; k < j; k++) s -= a_.coeff(i, k) * a_.coeff(j, k) * a_.coeff(k, k);
                a_.coeffRef(i, j) = s / d;
            }
        }
        return *this;
    }
    bool isPositive() const { for (int i = 0; i < a_.rows(); i++) if (!(a_.coeff(i, i) > T(0))) return false; return ok_; }
    bool isNegative() const { for (int i = 0; i < a_.rows(); i++) if (!(a_.coeff(i, i) < T(0))) return false; return ok_; }
    ComputationInfo info() const { return ok_ ? Success : NumericalIssue; }
    template <class D> typename internal::plain_of<D>::type solve(const MatrixBase<D>& b) const
    {
        typename internal::plain_of<D>::type x = b.eval();
        const int n = a_.rows();
        for (int c = 0; c < x.cols(); c++) {
            for (int i = 0; i < n; i++) { T s = x.coeff(i, c); for (int k = 0; k < i; k++) s -= a_.coeff(i, k) * x.coeff(k, c); x.coeffRef(i, c) = s; }
            for (int i = 0; i < n; i++) x.coeffRef(i, c) /= a_.coeff(i, i);
            for (int i = n - 1; i >= 0; i--) { T s = x.coeff(i, c); for (int k = i + 1; k < n; k++) s -= a_.coeff(k, i) * x.coeff(k, c); x.coeffRef(i, c) = s; }
        }
        return x;
    }
    Matrix<T, traits<M>::Rows, 1> vectorD() const { Matrix<T, traits<M>::Rows, 1> d; d.resize(a_.rows(), 1); for (int i = 0; i < a_.rows(); i++) d.coeffRef(i, 0) = a_.coeff(i, i); return d; }
};
template <class M> class LLT {
    typedef typename traits<M>::Scalar T;
    M l_;
    bool ok_ = false;
public:
    LLT() { }
    template <class D> explicit LLT(const MatrixBase<D>& a) { compute(a); }
    template <class D> LLT& compute(const MatrixBase<D>& a)
    {
        l_.assign(a.derived());
        const int n = l_.rows();
        ok_ = true;
        for (int j = 0; j < n; j++) {
            T d = l_.coeff(j, j);
            for (int k = 0; k < j; k++) d -= l_.coeff(j, k) * l_.coeff(j, k);
            if (!(d > T(0))) { ok_ = false; d = T(1); }
            d = std::sqrt(d);
            l_.coeffRef(j, j) = d;
            for (int i = j + 1; i < n; i++) {
                T s = l_.coeff(i, j);
                for (int k = 0; k < j; k++) s -= l_.coeff(i, k) * l_.coeff(j, k);
                l_.coeffRef(i, j) = s / d;
            }
            for (int i = 0; i < j; i++) l_.coeffRef(i, j) = T(0);
        }
        return *this;
    }
    ComputationInfo info() const { return ok_ ? Success : NumericalIssue; }
    M matrixL() const { return l_; }
    M matrixLLT() const { return l_; }
    template <class D> typename internal::plain_of<D>::type solve(const MatrixBase<D>& b) const
    {
        typename internal::plain_of<D>::type x = b.eval();
        const int n = l_.rows();
        for (int c = 0; c < x.cols(); c++) {
            for (int i = 0; i < n; i++) { T s = x.coeff(i, c); for (int k = 0; k < i; k++) s -= l_.coeff(i, k) * x.coeff(k, c); x.coeffRef(i, c) = s / l_.coeff(i, i); }
            for (int i = n - 1; i >= 0; i--) { T s = x.coeff(i, c); for (int k = i + 1; k < n; k++) s -= l_.coeff(k, i) * x.coeff(k, c); x.coeffRef(i, c) = s / l_.coeff(i, i); }
        }
        return x;
    }
};
/* LU with partial (row) pivoting, unblocked right-looking form as Eigen's partial_lu_impl::unblocked_lu */
template <class M> class PartialPivLU {
    typedef typename traits<M>::Scalar T;
    M lu_;
    std::vector<int> piv_;
public:
    PartialPivLU() { }
    template <class D> explicit PartialPivLU(const MatrixBase<D>& a) { compute(a); }
    template <class D> PartialPivLU& compute(const MatrixBase<D>& a)
    {
        lu_.assign(a.derived());
        const int n = lu_.rows();
        piv_.assign((size_t)n, 0);
        for (int k = 0; k < n; k++) {
            int p = k;
            for (int r = k + 1; r < n; r++) if (std::abs(lu_.coeff(r, k)) > std::abs(lu_.coeff(p, k))) p = r;
            piv_[(size_t)k] = p;
            if (p != k) for (int c = 0; c < n; c++) std::swap(lu_.coeffRef(k, c), lu_.coeffRef(p, c));
            if (lu_.coeff(k, k) != T(0)) for (int r = k + 1; r < n; r++) lu_.coeffRef(r, k) /= lu_.coeff(k, k);
            for (int c = k + 1; c < n; c++) for (int r = k + 1; r < n; r++) lu_.coeffRef(r, c) -= lu_.coeff(r, k) * lu_.coeff(k, c);
        }
        return *this;
    }
    template <class D> typename internal::plain_of<D>::type solve(const MatrixBase<D>& b) const
    {
        typename internal::plain_of<D>::type x = b.eval();
        const int n = lu_.rows();
        for (int c = 0; c < x.cols(); c++) {
            for (int k = 0; k < n; k++) if (piv_[(size_t)k] != k) std::swap(x.coeffRef(k, c), x.coeffRef(piv_[(size_t)k], c));
            for (int i = 0; i < n; i++) { T s = x.coeff(i, c); for (int k = 0; k < i; k++) s -= lu_.coeff(i, k) * x.coeff(k, c); x.coeffRef(i, c) = s; }
            for (int i = n - 1; i >= 0; i--) { T s = x.coeff(i, c); for (int k = i + 1; k < n; k++) s -= lu_.coeff(i, k) * x.coeff(k, c); x.coeffRef(i, c) = s / lu_.coeff(i, i); }
        }
        return x;
    }
    M inverse() const { M I; I.resize(lu_.rows(), lu_.rows()); I.setIdentity(); return solve(I); }
    T determinant() const { T d = 1; for (int k = 0; k < lu_.rows(); k++) { d *= lu_.coeff(k, k); if (piv_[(size_t)k] != k) d = -d; } return d; }
};
template <class D> PartialPivLU<typename internal::plain_of<D>::type> MatrixBase<D>::lu() const { return PartialPivLU<PlainObject>(*this); }
template <class D> PartialPivLU<typename internal::plain_of<D>::type> MatrixBase<D>::partialPivLu() const { return PartialPivLU<PlainObject>(*this); }
template <class D> LDLT<typename internal::plain_of<D>::type> MatrixBase<D>::ldlt() const { return LDLT<PlainObject>(*this); }
template <class D> LLT<typename internal::plain_of<D>::type> MatrixBase<D>::llt() const { return LLT<PlainObject>(*this); }

namespace internal {
/* Gauss-Jordan with partial pivoting (Eigen: cofactors up to 4x4, PartialPivLU above; same result up to rounding) */
template <class M> M gj_inverse(const M& a, typename traits<M>::Scalar* det_out)
{
    typedef typename traits<M>::Scalar T;
    const int n = a.rows();
    M w = a, inv;
    inv.resize(n, n);
    inv.setIdentity();
    T det = 1;
    for (int c = 0; c < n; c++) {
        int p = c;
        for (int r = c + 1; r < n; r++) if (std::abs(w.coeff(r, c)) > std::abs(w.coeff(p, c))) p = r;
        if (p != c) {
            for (int k = 0; k < n; k++) { std::swap(w.coeffRef(p, k), w.coeffRef(c, k)); std::swap(inv.coeffRef(p, k), inv.coeffRef(c, k)); }
            det = -det;
        }
        const T d = w.coeff(c, c);
        det *= d;
        const T id = T(1) / d;
        for (int k = 0; k < n; k++) { w.coeffRef(c, k) *= id; inv.coeffRef(c, k) *= id; }
        for (int r = 0; r < n; r++) {
            if (r == c) continue;
            const T f = w.coeff(r, c);
            if (f == T(0)) continue;
            for (int k = 0; k < n; k++) { w.coeffRef(r, k) -= f * w.coeff(c, k); inv.coeffRef(r, k) -= f * inv.coeff(c, k); }
        }
    }
    if (det_out) *det_out = det;
    return inv;
}
} // namespace internal
template <class D> typename MatrixBase<D>::PlainObject MatrixBase<D>::inverse() const
{
    const PlainObject a = eval();
    const int n = a.rows();
    if (n == 1) { PlainObject o = a; o.coeffRef(0, 0) = Scalar(1) / a.coeff(0, 0); return o; }
    if (n == 2) {
        PlainObject o = a;
        const Scalar id = Scalar(1) / (a.coeff(0, 0) * a.coeff(1, 1) - a.coeff(1, 0) * a.coeff(0, 1));
        o.coeffRef(0, 0) = a.coeff(1, 1) * id; o.coeffRef(1, 0) = -a.coeff(1, 0) * id;
        o.coeffRef(0, 1) = -a.coeff(0, 1) * id; o.coeffRef(1, 1) = a.coeff(0, 0) * id;
        return o;
    }
    if (n == 3) { /* cofactor form, as Eigen's compute_inverse<Matrix3> */
        PlainObject o = a;
        auto cof = [&](int i, int j) {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            return a.coeff(i1, j1) * a.coeff(i2, j2) - a.coeff(i1, j2) * a.coeff(i2, j1);
        };
        const Scalar c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
        const Scalar id = Scalar(1) / (c00 * a.coeff(0, 0) + c10 * a.coeff(1, 0) + c20 * a.coeff(2, 0));
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o.coeffRef(j, i) = cof(i, j) * id;
        return o;
    }
    return internal::gj_inverse(a, (Scalar*)nullptr);
}
template <class D> typename MatrixBase<D>::Scalar MatrixBase<D>::determinant() const
{
    const PlainObject a = eval();
    const int n = a.rows();
    if (n == 1) return a.coeff(0, 0);
    if (n == 2) return a.coeff(0, 0) * a.coeff(1, 1) - a.coeff(1, 0) * a.coeff(0, 1);
    if (n == 3)
        return a.coeff(0, 0) * (a.coeff(1, 1) * a.coeff(2, 2) - a.coeff(1, 2) * a.coeff(2, 1))
             - a.coeff(0, 1) * (a.coeff(1, 0) * a.coeff(2, 2) - a.coeff(1, 2) * a.coeff(2, 0))
             + a.coeff(0, 2) * (a.coeff(1, 0) * a.coeff(2, 1) - a.coeff(1, 1) * a.coeff(2, 0));
    Scalar det = 0;
    internal::gj_inverse(a, &det);
    return det;
}

/* symmetric eigenvalues by cyclic Jacobi rotations (only g2o's verifyInformationMatrices asks for them) */
template <class M> class SelfAdjointEigenSolver {
    typedef typename traits<M>::Scalar T;
    Matrix<T, Dynamic, 1> ev_;
public:
    SelfAdjointEigenSolver() { }
    template <class D> SelfAdjointEigenSolver& compute(const MatrixBase<D>& a_in, int = 0)
    {
        Matrix<T, Dynamic, Dynamic> a;
        a.assign(a_in.derived());
        const int n = a.rows();
        for (int sweep = 0; sweep < 64; sweep++) {
            T off = 0;
            for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) off += a.coeff(p, q) * a.coeff(p, q);
            if (off < T(1e-300)) break;
            for (int p = 0; p < n; p++)
                for (int q = p + 1; q < n; q++) {
                    if (a.coeff(p, q) == T(0)) continue;
                    const T th = (a.coeff(q, q) - a.coeff(p, p)) / (2 * a.coeff(p, q));
                    const T t = (th >= 0 ? T(1) : T(-1)) / (std::abs(th) + std::sqrt(th * th + 1));
                    const T c = T(1) / std::sqrt(t * t + 1), s = t * c;
                    for (int k = 0; k < n; k++) { const T akp = a.coeff(k, p), akq = a.coeff(k, q); a.coeffRef(k, p) = c * akp - s * akq; a.coeffRef(k, q) = s * akp + c * akq; }
                    for (int k = 0; k < n; k++) { const T apk = a.coeff(p, k), aqk = a.coeff(q, k); a.coeffRef(p, k) = c * apk - s * aqk; a.coeffRef(q, k) = s * apk + c * aqk; }
                }
        }
        ev_.resize(n, 1);
        for (int i = 0; i < n; i++) ev_.coeffRef(i, 0) = a.coeff(i, i);
        std::sort(ev_.data(), ev_.data() + n);
        return *this;
    }
    const Matrix<T, Dynamic, 1>& eigenvalues() const { return ev_; }
    ComputationInfo info() const { return Success; }
};

/* ---- typedefs ---- */
#define DVM_MINI_EIGEN_TYPEDEFS(T, S)                          \
    typedef Matrix<T, 2, 2> Matrix2##S; typedef Matrix<T, 3, 3> Matrix3##S; typedef Matrix<T, 4, 4> Matrix4##S; \
    typedef Matrix<T, Dynamic, Dynamic> MatrixX##S;            \
    typedef Matrix<T, 2, 1> Vector2##S; typedef Matrix<T, 3, 1> Vector3##S; typedef Matrix<T, 4, 1> Vector4##S; \
    typedef Matrix<T, Dynamic, 1> VectorX##S; typedef Matrix<T, 1, Dynamic> RowVectorX##S; \
    typedef Matrix<T, 1, 2> RowVector2##S; typedef Matrix<T, 1, 3> RowVector3##S; typedef Matrix<T, 1, 4> RowVector4##S;
DVM_MINI_EIGEN_TYPEDEFS(double, d)
DVM_MINI_EIGEN_TYPEDEFS(float, f)
DVM_MINI_EIGEN_TYPEDEFS(int, i)
#undef DVM_MINI_EIGEN_TYPEDEFS

/* ---- geometry ---- */
template <class T> class AngleAxis;
template <class T, int Options = 0> class Quaternion {
    Matrix<T, 4, 1> c_;   /* x, y, z, w */
public:
    typedef T Scalar;
    typedef Matrix<T, 3, 1> Vector3;
    typedef Matrix<T, 3, 3> Matrix3;
    Quaternion() { }
    Quaternion(T w, T x, T y, T z) { c_.coeffRef(0) = x; c_.coeffRef(1) = y; c_.coeffRef(2) = z; c_.coeffRef(3) = w; }
    Quaternion(const Quaternion& o) : c_(o.c_) { }
    template <class D> explicit Quaternion(const MatrixBase<D>& m) { if (m.rows() == 3 && m.cols() == 3) fromMatrix(m.derived()); else for (int i = 0; i < 4; i++) c_.coeffRef(i) = m.coeff(i); }
    explicit Quaternion(const T* d) { for (int i = 0; i < 4; i++) c_.coeffRef(i) = d[i]; }
    explicit Quaternion(const AngleAxis<T>& aa);
    Quaternion& operator=(const Quaternion& o) { c_ = o.c_; return *this; }
    template <class D> Quaternion& operator=(const MatrixBase<D>& m) { fromMatrix(m.derived()); return *this; }
    static Quaternion Identity() { return Quaternion(1, 0, 0, 0); }
    Quaternion& setIdentity() { *this = Identity(); return *this; }
    T x() const { return c_.coeff(0); } T y() const { return c_.coeff(1); } T z() const { return c_.coeff(2); } T w() const { return c_.coeff(3); }
    T& x() { return c_.coeffRef(0); } T& y() { return c_.coeffRef(1); } T& z() { return c_.coeffRef(2); } T& w() { return c_.coeffRef(3); }
    const Matrix<T, 4, 1>& coeffs() const { return c_; }
    Matrix<T, 4, 1>& coeffs() { return c_; }
    Vector3 vec() const { return Vector3(x(), y(), z()); }
    T squaredNorm() const { return c_.squaredNorm(); }
    T norm() const { return c_.norm(); }
    void normalize() { c_ /= c_.norm(); }   /* Eigen: m_coeffs /= norm() */
    Quaternion normalized() const { Quaternion q(*this); q.normalize(); return q; }
    Quaternion conjugate() const { return Quaternion(w(), -x(), -y(), -z()); }
    Quaternion inverse() const
    {
        const T n2 = squaredNorm();
        if (n2 > T(0)) { Quaternion q = conjugate(); q.c_ /= n2; return q; }
        Quaternion q; q.c_.setZero(); return q;
    }
    template <class U> Quaternion<U> cast() const { return Quaternion<U>((U)w(), (U)x(), (U)y(), (U)z()); }
    /* Eigen/src/Geometry/Quaternion.h: quat_product */
    Quaternion operator*(const Quaternion& b) const
    {
        const Quaternion& a = *this;
        return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                          a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                          a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                          a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
    }
    Quaternion& operator*=(const Quaternion& b) { *this = *this * b; return *this; }
    /* _transformVector: v + w * 2(qv x v) + qv x 2(qv x v) */
    template <class D> Vector3 operator*(const MatrixBase<D>& v_in) const
    {
        const Vector3 v = v_in.eval();
        Vector3 uv = vec().cross(v);
        uv += uv;
        const Vector3 c = vec().cross(uv);
        return Vector3(v.coeff(0) + w() * uv.coeff(0) + c.coeff(0), v.coeff(1) + w() * uv.coeff(1) + c.coeff(1), v.coeff(2) + w() * uv.coeff(2) + c.coeff(2));
    }
    Vector3 _transformVector(const Vector3& v) const { return (*this) * v; }
    Matrix3 toRotationMatrix() const
    {
        Matrix3 res;
        const T tx = T(2) * x(), ty = T(2) * y(), tz = T(2) * z();
        const T twx = tx * w(), twy = ty * w(), twz = tz * w();
        const T txx = tx * x(), txy = ty * x(), txz = tz * x();
        const T tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
        res.coeffRef(0, 0) = T(1) - (tyy + tzz); res.coeffRef(0, 1) = txy - twz; res.coeffRef(0, 2) = txz + twy;
        res.coeffRef(1, 0) = txy + twz; res.coeffRef(1, 1) = T(1) - (txx + tzz); res.coeffRef(1, 2) = tyz - twx;
        res.coeffRef(2, 0) = txz - twy; res.coeffRef(2, 1) = tyz + twx; res.coeffRef(2, 2) = T(1) - (txx + tyy);
        return res;
    }
    Matrix3 matrix() const { return toRotationMatrix(); }
    T angularDistance(const Quaternion& o) const
    {
        const Quaternion d = (*this) * o.conjugate();
        return T(2) * std::atan2(d.vec().norm(), std::abs(d.w()));
    }
    bool isApprox(const Quaternion& o, T prec = T(1e-12)) const { return c_.isApprox(o.c_, prec); }
private:
    /* Eigen/src/Geometry/Quaternion.h: quaternionbase_assign_impl<Other,3,3> */
    template <class M> void fromMatrix(const M& mat)
    {
        T t = mat.coeff(0, 0) + mat.coeff(1, 1) + mat.coeff(2, 2);
        if (t > T(0)) {
            t = std::sqrt(t + T(1.0));
            w() = T(0.5) * t;
            t = T(0.5) / t;
            x() = (mat.coeff(2, 1) - mat.coeff(1, 2)) * t;
            y() = (mat.coeff(0, 2) - mat.coeff(2, 0)) * t;
            z() = (mat.coeff(1, 0) - mat.coeff(0, 1)) * t;
        } else {
            int i = 0;
            if (mat.coeff(1, 1) > mat.coeff(0, 0)) i = 1;
            if (mat.coeff(2, 2) > mat.coeff(i, i)) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(mat.coeff(i, i) - mat.coeff(j, j) - mat.coeff(k, k) + T(1.0));
            c_.coeffRef(i) = T(0.5) * t;
            t = T(0.5) / t;
            w() = (mat.coeff(k, j) - mat.coeff(j, k)) * t;
            c_.coeffRef(j) = (mat.coeff(j, i) + mat.coeff(i, j)) * t;
            c_.coeffRef(k) = (mat.coeff(k, i) + mat.coeff(i, k)) * t;
        }
    }
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;
template <class T, int O> std::ostream& operator<<(std::ostream& os, const Quaternion<T, O>& q) { return os << q.x() << "i + " << q.y() << "j + " << q.z() << "k + " << q.w(); }

template <class T> class AngleAxis {
    Matrix<T, 3, 1> axis_;
    T angle_ = 0;
public:
    AngleAxis() { }
    template <class D> AngleAxis(T angle, const MatrixBase<D>& axis) : axis_(axis), angle_(angle) { }
    explicit AngleAxis(const Quaternion<T>& q)
    {
        T n = q.vec().norm();
        if (n < std::numeric_limits<T>::epsilon()) n = q.vec().squaredNorm() > 0 ? std::sqrt(q.vec().squaredNorm()) : T(0);
        if (n != T(0)) { angle_ = T(2) * std::atan2(n, std::abs(q.w())); if (q.w() < T(0)) n = -n; axis_ = q.vec() / n; }
        else { angle_ = T(0); axis_ = Matrix<T, 3, 1>(1, 0, 0); }
    }
    template <class D> explicit AngleAxis(const MatrixBase<D>& R) { *this = AngleAxis(Quaternion<T>(R)); }
    T angle() const { return angle_; }
    T& angle() { return angle_; }
    const Matrix<T, 3, 1>& axis() const { return axis_; }
    Matrix<T, 3, 1>& axis() { return axis_; }
    Matrix<T, 3, 3> toRotationMatrix() const { return Quaternion<T>(*this).toRotationMatrix(); }
    Matrix<T, 3, 3> matrix() const { return toRotationMatrix(); }
};
template <class T, int O> Quaternion<T, O>::Quaternion(const AngleAxis<T>& aa)
{
    const T ha = T(0.5) * aa.angle();
    w() = std::cos(ha);
    const Matrix<T, 3, 1> v = aa.axis() * std::sin(ha);
    x() = v.coeff(0); y() = v.coeff(1); z() = v.coeff(2);
}
typedef AngleAxis<double> AngleAxisd;
typedef AngleAxis<float> AngleAxisf;

/* Transform<T, Dim, Mode>: a (Dim+1)^2 homogeneous matrix with the accessors g2o's headers name */
template <class T, int Dim, int Mode, int Options = 0> class Transform {
    Matrix<T, Dim + 1, Dim + 1> m_;
public:
    typedef Matrix<T, Dim, Dim> LinearMatrixType;
    typedef Matrix<T, Dim, 1> VectorType;
    Transform() { m_.setIdentity(); }
    template <class D> Transform(const MatrixBase<D>& m) { m_.setIdentity(); if (m.rows() == Dim + 1) m_ = m; else m_.template block<Dim, Dim>(0, 0) = m; }
    Transform(const Quaternion<T>& q) { m_.setIdentity(); m_.template block<Dim, Dim>(0, 0) = q.toRotationMatrix(); }
    template <class D> Transform& operator=(const MatrixBase<D>& m) { *this = Transform(m); return *this; }
    Transform& operator=(const Quaternion<T>& q) { *this = Transform(q); return *this; }
    static Transform Identity() { return Transform(); }
    void setIdentity() { m_.setIdentity(); }
    Block<Matrix<T, Dim + 1, Dim + 1>, Dim, Dim> linear() { return m_.template block<Dim, Dim>(0, 0); }
    LinearMatrixType linear() const { LinearMatrixType l = m_.template block<Dim, Dim>(0, 0); return l; }
    Block<Matrix<T, Dim + 1, Dim + 1>, Dim, Dim> rotation_ref() { return linear(); }
    LinearMatrixType rotation() const { return linear(); }
    Block<Matrix<T, Dim + 1, Dim + 1>, Dim, 1> translation() { return m_.template block<Dim, 1>(0, Dim); }
    VectorType translation() const { VectorType t = m_.template block<Dim, 1>(0, Dim); return t; }
    Matrix<T, Dim + 1, Dim + 1>& matrix() { return m_; }
    const Matrix<T, Dim + 1, Dim + 1>& matrix() const { return m_; }
    T operator()(int i, int j) const { return m_.coeff(i, j); }
    T& operator()(int i, int j) { return m_.coeffRef(i, j); }
    Transform operator*(const Transform& o) const { Transform r; r.m_ = m_ * o.m_; return r; }
    template <class D> VectorType operator*(const MatrixBase<D>& v) const { VectorType r = linear() * v + translation(); return r; }
    Transform inverse(int = 0) const
    {
        Transform r;
        if (Mode == Isometry) {
            const LinearMatrixType Rt = linear().transpose();
            r.m_.template block<Dim, Dim>(0, 0) = Rt;
            r.m_.template block<Dim, 1>(0, Dim) = -(Rt * translation());
        } else r.m_ = m_.inverse();
        return r;
    }
    Transform& translate(const VectorType& v) { m_.template block<Dim, 1>(0, Dim) = translation() + linear() * v; return *this; }
    Transform& pretranslate(const VectorType& v) { m_.template block<Dim, 1>(0, Dim) = translation() + v; return *this; }
};
typedef Transform<double, 3, Isometry> Isometry3d;
typedef Transform<double, 2, Isometry> Isometry2d;
typedef Transform<double, 3, Affine> Affine3d;
typedef Transform<double, 2, Affine> Affine2d;
typedef Transform<float, 3, Isometry> Isometry3f;

/* ---- the sparse interface g2o's LinearSolverEigen names: a dense symmetric matrix behind Eigen::SparseMatrix's calls,
 * factored by LDL^T without reordering (the AMD ordering of SimplicialLDLT changes rounding only) ---- */
template <class T, class I = int> struct Triplet {
    I r, c; T v;
    Triplet() : r(0), c(0), v(0) { }
    Triplet(I r_, I c_, T v_ = T(0)) : r(r_), c(c_), v(v_) { }
    I row() const { return r; } I col() const { return c; } T value() const { return v; }
};
template <int SizeAtCompileTime, int MaxSize = SizeAtCompileTime, class I = int> class PermutationMatrix {
    Matrix<I, Dynamic, 1> idx_;
public:
    PermutationMatrix() { }
    void resize(int n) { idx_.resize(n, 1); }
    int size() const { return idx_.rows(); }
    Matrix<I, Dynamic, 1>& indices() { return idx_; }
    const Matrix<I, Dynamic, 1>& indices() const { return idx_; }
    PermutationMatrix inverse() const { PermutationMatrix p; p.resize(size()); for (int i = 0; i < size(); i++) p.idx_.coeffRef(idx_.coeff(i)) = i; return p; }
};
} // namespace Eigen
#endif
