// ORACLE / reference pin (test infrastructure only).
//
// A small stand-in for the part of Eigen 3.3 that the reference's hot-path text uses, so that text
// extracted VERBATIM from /root/reference (oracle/ref/extract.sh) compiles in this image, which has
// no Eigen. Everything is evaluated eagerly (no expression templates); the arithmetic that Eigen
// fixes is kept where it can change a rounding:
//   * dot / squaredNorm / norm of a fixed 3-vector reduce as (e0+e1)+e2 (Eigen 3.3 Redux.h,
//     LinearVectorizedTraversal + CompleteUnrolling with SSE2 Packet2d: predux of the first packet,
//     then the scalar tail); longer reductions are sequential,
//   * matrix products accumulate k = 0, 1, 2, ... in order (coefficient-based lazy product,
//     packet path of ProductEvaluators.h),
//   * normalize() divides by sqrt(squaredNorm()) and leaves a zero vector alone (Dot.h),
//   * operator/=(scalar) and operator/(scalar) divide, they do not multiply by a reciprocal,
//   * Quaternion: Eigen 3.3 Quaternion.h formulas (generic product, _transformVector, slerp,
//     rotation-matrix conversions).
// Decompositions (SelfAdjointEigenSolver, ColPivHouseholderQR, JacobiSVD, LLT, inverse) are
// restatements of the published algorithms, not Eigen's code: results agree to rounding, not bit
// for bit. This is stated wherever a parity claim depends on it.
#ifndef MML_REF_EIGEN_H
#define MML_REF_EIGEN_H
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <limits>
#include <type_traits>
#include <utility>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {

constexpr int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 1 };
enum { ComputeFullU = 0x04, ComputeThinU = 0x08, ComputeFullV = 0x10, ComputeThinV = 0x20 };

template <class T, int R, int C, int Opt = 0> class Matrix;
template <class X, int BR, int BC> class Block;
template <class M> class Map;
template <class D> struct traits;

namespace detail {
using std::sqrt;
using std::abs;
template <class T> struct is_scalar_like : std::is_arithmetic<T> {};
constexpr int pick(int a, int b) { return a != Dynamic ? a : b; }

template <class T, int R, int C> struct Storage {
  T d[R * C > 0 ? R * C : 1];
  Storage() { for (int i = 0; i < R * C; i++) d[i] = T(); }
  int rows() const { return R; }
  int cols() const { return C; }
  void resize(int, int) {}
  T* data() { return d; }
  const T* data() const { return d; }
};
template <class T, int R, int C, bool Dyn = (R == Dynamic || C == Dynamic)> struct StorageSel { using type = Storage<T, R, C>; };
template <class T, int R, int C> struct DynStorage {
  std::vector<T> d;
  int r = (R == Dynamic ? 0 : R), c = (C == Dynamic ? 0 : C);
  DynStorage() { d.assign((size_t)r * c, T()); }
  int rows() const { return r; }
  int cols() const { return c; }
  void resize(int rr, int cc) { r = rr; c = cc; d.assign((size_t)r * c, T()); }
  T* data() { return d.data(); }
  const T* data() const { return d.data(); }
};
template <class T, int R, int C> struct StorageSel<T, R, C, true> { using type = DynStorage<T, R, C>; };
}  // namespace detail

template <class D> class MatrixBase;
template <class M> class ColPivHouseholderQR;
template <class T> struct ArrayX;

template <class D> class CommaInit {
  D& m; int k;
 public:
  CommaInit(D& m_, const typename traits<D>::Scalar& v) : m(m_), k(0) { put(v); }
  void put(const typename traits<D>::Scalar& v) { int c = m.cols(); m.coeffRef(k / c, k % c) = v; k++; }
  CommaInit& operator,(const typename traits<D>::Scalar& v) { put(v); return *this; }
};

template <class D> class MatrixBase {
 public:
  using Scalar = typename traits<D>::Scalar;
  enum { RowsAtCompileTime = traits<D>::Rows, ColsAtCompileTime = traits<D>::Cols };
  using PlainObject = Matrix<Scalar, traits<D>::Rows, traits<D>::Cols>;
  const D& derived() const { return *static_cast<const D*>(this); }
  D& derived() { return *static_cast<D*>(this); }
  int rows() const { return derived().rows(); }
  int cols() const { return derived().cols(); }
  int size() const { return rows() * cols(); }
  // element access (vectors: linear index)
  Scalar operator()(int i, int j) const { return derived().coeff(i, j); }
  Scalar& operator()(int i, int j) { return derived().coeffRef(i, j); }
  Scalar lin(int i) const { return cols() == 1 ? derived().coeff(i, 0) : derived().coeff(0, i); }
  Scalar& linRef(int i) { return cols() == 1 ? derived().coeffRef(i, 0) : derived().coeffRef(0, i); }
  Scalar operator()(int i) const { return lin(i); }
  Scalar& operator()(int i) { return linRef(i); }
  Scalar operator[](int i) const { return lin(i); }
  Scalar& operator[](int i) { return linRef(i); }
  Scalar x() const { return lin(0); }  Scalar& x() { return linRef(0); }
  Scalar y() const { return lin(1); }  Scalar& y() { return linRef(1); }
  Scalar z() const { return lin(2); }  Scalar& z() { return linRef(2); }
  Scalar w() const { return lin(3); }  Scalar& w() { return linRef(3); }
  PlainObject eval() const {
    PlainObject r; r.resize(rows(), cols());
    for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) r.coeffRef(i, j) = derived().coeff(i, j);
    return r;
  }
  Matrix<Scalar, traits<D>::Cols, traits<D>::Rows> transpose() const {
    Matrix<Scalar, traits<D>::Cols, traits<D>::Rows> r; r.resize(cols(), rows());
    for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) r.coeffRef(j, i) = derived().coeff(i, j);
    return r;
  }
  template <class U> Matrix<U, traits<D>::Rows, traits<D>::Cols> cast() const {
    Matrix<U, traits<D>::Rows, traits<D>::Cols> r; r.resize(rows(), cols());
    for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) r.coeffRef(i, j) = U(derived().coeff(i, j));
    return r;
  }
  // reductions. Fixed-size 3: (e0+e1)+e2; otherwise sequential (see header comment).
  template <class O> Scalar dot(const MatrixBase<O>& o) const {
    int n = size();
    Scalar s = lin(0) * o.lin(0);
    for (int i = 1; i < n; i++) s = s + lin(i) * o.lin(i);
    return s;
  }
  Scalar squaredNorm() const {
    Scalar s = Scalar(0);
    bool first = true;
    for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) {
      Scalar v = derived().coeff(i, j);
      if (first) { s = v * v; first = false; } else s = s + v * v;
    }
    return s;
  }
  Scalar norm() const { using std::sqrt; return sqrt(squaredNorm()); }
  Scalar sum() const {
    Scalar s = Scalar(0); bool first = true;
    for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) {
      if (first) { s = derived().coeff(i, j); first = false; } else s = s + derived().coeff(i, j);
    }
    return s;
  }
  Scalar trace() const { Scalar s = derived().coeff(0, 0); for (int i = 1; i < rows(); i++) s = s + derived().coeff(i, i); return s; }
  void normalize() {
    Scalar z = squaredNorm();
    if (z > Scalar(0)) { using std::sqrt; Scalar s = sqrt(z); derived() /= s; }
  }
  PlainObject normalized() const { PlainObject r = eval(); r.normalize(); return r; }
  template <class O> Matrix<Scalar, 3, 1> cross(const MatrixBase<O>& o) const {
    Matrix<Scalar, 3, 1> r;
    r.coeffRef(0, 0) = lin(1) * o.lin(2) - lin(2) * o.lin(1);
    r.coeffRef(1, 0) = lin(2) * o.lin(0) - lin(0) * o.lin(2);
    r.coeffRef(2, 0) = lin(0) * o.lin(1) - lin(1) * o.lin(0);
    return r;
  }
  Scalar maxCoeff(int* ri = nullptr, int* ci = nullptr) const {
    Scalar m = derived().coeff(0, 0); int br = 0, bc = 0;
    for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++)
      if (derived().coeff(i, j) > m) { m = derived().coeff(i, j); br = i; bc = j; }
    if (ri) *ri = br; if (ci) *ci = bc;
    return m;
  }
  // setters
  D& setZero() { for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) derived().coeffRef(i, j) = Scalar(0); return derived(); }
  D& setOnes() { for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) derived().coeffRef(i, j) = Scalar(1); return derived(); }
  D& setConstant(const Scalar& v) { for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) derived().coeffRef(i, j) = v; return derived(); }
  D& setIdentity() { for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) derived().coeffRef(i, j) = Scalar(i == j ? 1 : 0); return derived(); }
  CommaInit<D> operator<<(const Scalar& v) { return CommaInit<D>(derived(), v); }
  // compound assignment
  template <class O> D& operator+=(const MatrixBase<O>& o) { auto t = o.eval(); for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) derived().coeffRef(i, j) = derived().coeff(i, j) + t.coeff(i, j); return derived(); }
  template <class O> D& operator-=(const MatrixBase<O>& o) { auto t = o.eval(); for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) derived().coeffRef(i, j) = derived().coeff(i, j) - t.coeff(i, j); return derived(); }
  D& operator*=(const Scalar& s) { for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) derived().coeffRef(i, j) = derived().coeff(i, j) * s; return derived(); }
  D& operator/=(const Scalar& s) { for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) derived().coeffRef(i, j) = derived().coeff(i, j) / s; return derived(); }
  template <class O> void applyOnTheLeft(const MatrixBase<O>& o);
  // blocks
  template <int BR, int BC> Block<D, BR, BC> block(int i, int j) { return Block<D, BR, BC>(derived(), i, j, BR, BC); }
  template <int BR, int BC> Block<const D, BR, BC> block(int i, int j) const { return Block<const D, BR, BC>(derived(), i, j, BR, BC); }
  Block<D, Dynamic, Dynamic> block(int i, int j, int r, int c) { return Block<D, Dynamic, Dynamic>(derived(), i, j, r, c); }
  Block<const D, Dynamic, Dynamic> block(int i, int j, int r, int c) const { return Block<const D, Dynamic, Dynamic>(derived(), i, j, r, c); }
  template <int N> Block<D, N, 1> segment(int i) { return Block<D, N, 1>(derived(), i, 0, N, 1); }
  template <int N> Block<const D, N, 1> segment(int i) const { return Block<const D, N, 1>(derived(), i, 0, N, 1); }
  Block<D, Dynamic, 1> segment(int i, int n) { return Block<D, Dynamic, 1>(derived(), i, 0, n, 1); }
  Block<const D, Dynamic, 1> segment(int i, int n) const { return Block<const D, Dynamic, 1>(derived(), i, 0, n, 1); }
  template <int N> Block<D, N, 1> head() { return segment<N>(0); }
  template <int N> Block<const D, N, 1> head() const { return segment<N>(0); }
  template <int N> Block<D, N, 1> tail() { return segment<N>(rows() - N); }
  template <int N> Block<const D, N, 1> tail() const { return segment<N>(rows() - N); }
  Block<D, Dynamic, Dynamic> topLeftCorner(int r, int c) { return block(0, 0, r, c); }
  Block<const D, Dynamic, Dynamic> topLeftCorner(int r, int c) const { return block(0, 0, r, c); }
  Block<D, Dynamic, Dynamic> topRightCorner(int r, int c) { return block(0, cols() - c, r, c); }
  Block<const D, Dynamic, Dynamic> topRightCorner(int r, int c) const { return block(0, cols() - c, r, c); }
  template <int BR, int BC> Block<D, BR, BC> topLeftCorner() { return block<BR, BC>(0, 0); }
  template <int BR, int BC> Block<const D, BR, BC> topLeftCorner() const { return block<BR, BC>(0, 0); }
  template <int BR, int BC> Block<D, BR, BC> topRightCorner() { return block<BR, BC>(0, cols() - BC); }
  template <int BR, int BC> Block<const D, BR, BC> topRightCorner() const { return block<BR, BC>(0, cols() - BC); }
  Block<D, traits<D>::Rows, 1> col(int j) { return Block<D, traits<D>::Rows, 1>(derived(), 0, j, rows(), 1); }
  Block<const D, traits<D>::Rows, 1> col(int j) const { return Block<const D, traits<D>::Rows, 1>(derived(), 0, j, rows(), 1); }
  Block<D, 1, traits<D>::Cols> row(int i) { return Block<D, 1, traits<D>::Cols>(derived(), i, 0, 1, cols()); }
  Block<const D, 1, traits<D>::Cols> row(int i) const { return Block<const D, 1, traits<D>::Cols>(derived(), i, 0, 1, cols()); }
  Block<D, traits<D>::Rows, Dynamic> leftCols(int n) { return Block<D, traits<D>::Rows, Dynamic>(derived(), 0, 0, rows(), n); }
  Block<D, traits<D>::Rows, Dynamic> middleCols(int j, int n) { return Block<D, traits<D>::Rows, Dynamic>(derived(), 0, j, rows(), n); }
  Block<const D, traits<D>::Rows, Dynamic> middleCols(int j, int n) const { return Block<const D, traits<D>::Rows, Dynamic>(derived(), 0, j, rows(), n); }
  // dense inverse (Gauss-Jordan with partial pivoting; restatement, see header)
  PlainObject inverse() const;
  ColPivHouseholderQR<PlainObject> colPivHouseholderQr() const;
  ArrayX<Scalar> array() const;
  Matrix<Scalar, Dynamic, Dynamic> asDiagonal() const;
  PlainObject cwiseSqrt() const { PlainObject r = eval(); using std::sqrt; for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) r.coeffRef(i, j) = sqrt(r.coeff(i, j)); return r; }
  PlainObject cwiseAbs() const { PlainObject r = eval(); using std::abs; for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) r.coeffRef(i, j) = abs(r.coeff(i, j)); return r; }
};

template <class T, int R, int C, int Opt> struct traits<Matrix<T, R, C, Opt>> { using Scalar = T; enum { Rows = R, Cols = C }; };

template <class T, int R, int C, int Opt>
class Matrix : public MatrixBase<Matrix<T, R, C, Opt>> {
  typename detail::StorageSel<T, R, C>::type s_;
  using Base = MatrixBase<Matrix<T, R, C, Opt>>;
 public:
  using Scalar = T;
  Matrix() {}
  // sizes (dynamic) or 2 coefficients (fixed 2-vector): only the dynamic reading is needed here
  Matrix(int r, int c) { if (R == Dynamic || C == Dynamic) s_.resize(r, c); else { s_.data()[0] = T(r); s_.data()[1] = T(c); } }
  explicit Matrix(int n) {
    if (R == Dynamic && C == 1) s_.resize(n, 1);
    else if (R == 1 && C == Dynamic) s_.resize(1, n);
    else s_.data()[0] = T(n);  // 1x1 from an int
  }
  template <class S, class = typename std::enable_if<std::is_floating_point<S>::value && R * C == 1, S>::type>
  explicit Matrix(const S& v) { s_.data()[0] = T(v); }
  Matrix(const T& a, const T& b, const T& c) { static_assert(R * C == 3, "3-vector"); s_.data()[0] = a; s_.data()[1] = b; s_.data()[2] = c; }
  Matrix(const T& a, const T& b, const T& c, const T& d) { static_assert(R * C == 4, "4-vector"); s_.data()[0] = a; s_.data()[1] = b; s_.data()[2] = c; s_.data()[3] = d; }
  Matrix(std::initializer_list<T> l) { int k = 0; for (const T& v : l) s_.data()[k++] = v; }
  template <class O> Matrix(const MatrixBase<O>& o) { *this = o; }
  template <class O> Matrix& operator=(const MatrixBase<O>& o) {
    // evaluate first: aliasing-safe
    std::vector<T> tmp((size_t)o.rows() * o.cols());
    int r = o.rows(), c = o.cols();
    for (int j = 0; j < c; j++) for (int i = 0; i < r; i++) tmp[(size_t)j * r + i] = o.derived().coeff(i, j);
    if ((R == Dynamic || C == Dynamic)) { if (s_.rows() != r || s_.cols() != c) s_.resize(r, c); }
    else if (R * C == r * c && (R != r)) { r = R; c = C; }  // vector <- transposed-shape vector
    for (int j = 0; j < c; j++) for (int i = 0; i < r; i++) coeffRef(i, j) = tmp[(size_t)j * r + i];
    return *this;
  }
  int rows() const { return s_.rows(); }
  int cols() const { return s_.cols(); }
  void resize(int r, int c) { s_.resize(r, c); }
  void resize(int n) { if (C == 1) s_.resize(n, 1); else s_.resize(1, n); }
  Matrix& setZero() { Base::setZero(); return *this; }
  Matrix& setZero(int r, int c) { s_.resize(r, c); Base::setZero(); return *this; }
  T coeff(int i, int j) const { return (Opt & RowMajor) ? s_.data()[(size_t)i * cols() + j] : s_.data()[(size_t)j * rows() + i]; }
  T& coeffRef(int i, int j) { return (Opt & RowMajor) ? s_.data()[(size_t)i * cols() + j] : s_.data()[(size_t)j * rows() + i]; }
  T* data() { return s_.data(); }
  const T* data() const { return s_.data(); }
  static Matrix Zero() { Matrix m; m.setZero(); return m; }
  static Matrix Zero(int r, int c) { Matrix m(r, c); m.setZero(); return m; }
  static Matrix Zero(int n) { Matrix m(n); m.setZero(); return m; }
  static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
  static Matrix Identity(int r, int c) { Matrix m(r, c); m.setIdentity(); return m; }
  static Matrix Ones() { Matrix m; m.setOnes(); return m; }
  static Matrix Constant(const T& v) { Matrix m; m.setConstant(v); return m; }
};

template <class X, int BR, int BC> struct traits<Block<X, BR, BC>> {
  using Scalar = typename traits<typename std::remove_const<X>::type>::Scalar; enum { Rows = BR, Cols = BC };
};
template <class X, int BR, int BC>
class Block : public MatrixBase<Block<X, BR, BC>> {
  X& x_; int i0, j0, r_, c_;
 public:
  using Scalar = typename traits<Block>::Scalar;
  Block(X& x, int i, int j, int r, int c) : x_(x), i0(i), j0(j), r_(r), c_(c) {}
  int rows() const { return r_; }
  int cols() const { return c_; }
  Scalar coeff(int i, int j) const { return x_.coeff(i0 + i, j0 + j); }
  Scalar& coeffRef(int i, int j) { return const_cast<typename std::remove_const<X>::type&>(x_).coeffRef(i0 + i, j0 + j); }
  template <class O> Block& operator=(const MatrixBase<O>& o) {
    auto t = o.eval();
    if (t.rows() == r_ && t.cols() == c_) { for (int j = 0; j < c_; j++) for (int i = 0; i < r_; i++) coeffRef(i, j) = t.coeff(i, j); }
    else { int k = 0; for (int j = 0; j < c_; j++) for (int i = 0; i < r_; i++, k++) coeffRef(i, j) = t.lin(k); }
    return *this;
  }
  Block& operator=(const Block& o) { return this->template operator=<Block>(o); }
};

template <class M> struct traits<Map<M>> {
  using Plain = typename std::remove_const<M>::type;
  using Scalar = typename traits<Plain>::Scalar; enum { Rows = traits<Plain>::Rows, Cols = traits<Plain>::Cols };
};
template <class T, int R, int C, int Opt> struct matrix_opt { enum { value = Opt }; };
template <class M> struct map_opt;
template <class T, int R, int C, int Opt> struct map_opt<Matrix<T, R, C, Opt>> { enum { value = Opt }; };
template <class M>
class Map : public MatrixBase<Map<M>> {
  using Plain = typename std::remove_const<M>::type;
  using T = typename traits<Plain>::Scalar;
  using Ptr = typename std::conditional<std::is_const<M>::value, const T*, T*>::type;
  Ptr p_; int r_, c_;
  enum { Opt = map_opt<Plain>::value };
 public:
  using Scalar = T;
  Map(Ptr p) : p_(p), r_(traits<Plain>::Rows), c_(traits<Plain>::Cols) {}
  Map(Ptr p, int n) : p_(p), r_(traits<Plain>::Cols == 1 ? n : 1), c_(traits<Plain>::Cols == 1 ? 1 : n) {}
  Map(Ptr p, int r, int c) : p_(p), r_(r), c_(c) {}
  int rows() const { return r_; }
  int cols() const { return c_; }
  T coeff(int i, int j) const { return (Opt & RowMajor) ? p_[(size_t)i * c_ + j] : p_[(size_t)j * r_ + i]; }
  T& coeffRef(int i, int j) { return const_cast<T*>(p_)[(Opt & RowMajor) ? (size_t)i * c_ + j : (size_t)j * r_ + i]; }
  template <class O> Map& operator=(const MatrixBase<O>& o) {
    auto t = o.eval();
    for (int j = 0; j < c_; j++) for (int i = 0; i < r_; i++) coeffRef(i, j) = t.coeff(i, j);
    return *this;
  }
  Map& operator=(const Map& o) { return this->template operator=<Map>(o); }
};

// ---- free operators (eager) ---------------------------------------------------------------
#define MML_RES(A, B) Matrix<typename traits<A>::Scalar, detail::pick(traits<A>::Rows, traits<B>::Rows), detail::pick(traits<A>::Cols, traits<B>::Cols)>
template <class A, class B> MML_RES(A, B) operator+(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  MML_RES(A, B) r; r.resize(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); j++) for (int i = 0; i < a.rows(); i++) r.coeffRef(i, j) = a.derived().coeff(i, j) + b.derived().coeff(i, j);
  return r;
}
template <class A, class B> MML_RES(A, B) operator-(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  MML_RES(A, B) r; r.resize(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); j++) for (int i = 0; i < a.rows(); i++) r.coeffRef(i, j) = a.derived().coeff(i, j) - b.derived().coeff(i, j);
  return r;
}
#undef MML_RES
template <class A> typename MatrixBase<A>::PlainObject operator-(const MatrixBase<A>& a) {
  typename MatrixBase<A>::PlainObject r; r.resize(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); j++) for (int i = 0; i < a.rows(); i++) r.coeffRef(i, j) = -a.derived().coeff(i, j);
  return r;
}
template <class A, class B>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> operator*(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  using S = typename traits<A>::Scalar;
  Matrix<S, traits<A>::Rows, traits<B>::Cols> r; r.resize(a.rows(), b.cols());
  const int K = a.cols();
  for (int j = 0; j < b.cols(); j++) for (int i = 0; i < a.rows(); i++) {
    S s = a.derived().coeff(i, 0) * b.derived().coeff(0, j);
    for (int k = 1; k < K; k++) s = s + a.derived().coeff(i, k) * b.derived().coeff(k, j);
    r.coeffRef(i, j) = s;
  }
  return r;
}
template <class S, class A> struct scalar_ok : std::integral_constant<bool, std::is_arithmetic<S>::value || std::is_same<S, typename traits<A>::Scalar>::value> {};
template <class S, class A, class = typename std::enable_if<scalar_ok<S, A>::value>::type>
typename MatrixBase<A>::PlainObject operator*(const S& s, const MatrixBase<A>& a) {
  typename MatrixBase<A>::PlainObject r; r.resize(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); j++) for (int i = 0; i < a.rows(); i++) r.coeffRef(i, j) = s * a.derived().coeff(i, j);
  return r;
}
template <class S, class A, class = typename std::enable_if<scalar_ok<S, A>::value>::type>
typename MatrixBase<A>::PlainObject operator*(const MatrixBase<A>& a, const S& s) {
  typename MatrixBase<A>::PlainObject r; r.resize(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); j++) for (int i = 0; i < a.rows(); i++) r.coeffRef(i, j) = a.derived().coeff(i, j) * s;
  return r;
}
template <class S, class A, class = typename std::enable_if<scalar_ok<S, A>::value>::type>
typename MatrixBase<A>::PlainObject operator/(const MatrixBase<A>& a, const S& s) {
  typename MatrixBase<A>::PlainObject r; r.resize(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); j++) for (int i = 0; i < a.rows(); i++) r.coeffRef(i, j) = a.derived().coeff(i, j) / s;
  return r;
}
template <class D> template <class O> void MatrixBase<D>::applyOnTheLeft(const MatrixBase<O>& o) {
  auto t = (o * (*this)).eval();
  for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) derived().coeffRef(i, j) = t.coeff(i, j);
}
template <class D> typename MatrixBase<D>::PlainObject MatrixBase<D>::inverse() const {
  const int n = rows();
  std::vector<Scalar> a((size_t)n * 2 * n);
  auto A = [&](int i, int j) -> Scalar& { return a[(size_t)i * 2 * n + j]; };
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { A(i, j) = derived().coeff(i, j); A(i, n + j) = Scalar(i == j ? 1 : 0); }
  for (int c = 0; c < n; c++) {
    int p = c; using std::abs;
    for (int i = c + 1; i < n; i++) if (abs(A(i, c)) > abs(A(p, c))) p = i;
    if (p != c) for (int j = 0; j < 2 * n; j++) std::swap(A(p, j), A(c, j));
    Scalar d = A(c, c);
    for (int j = 0; j < 2 * n; j++) A(c, j) = A(c, j) / d;
    for (int i = 0; i < n; i++) if (i != c) { Scalar f = A(i, c); if (f == Scalar(0)) continue; for (int j = 0; j < 2 * n; j++) A(i, j) = A(i, j) - f * A(c, j); }
  }
  PlainObject r; r.resize(n, n);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) r.coeffRef(i, j) = A(i, n + j);
  return r;
}

// coefficient-wise views used by the marginalisation text: (v.array() > eps).select(a, 0), .inverse()
template <class T> struct BoolArrayX {
  std::vector<char> b;
  Matrix<T, Dynamic, 1> select(const ArrayX<T>& a, const T& other) const;
};
template <class T> struct ArrayX {
  std::vector<T> v;
  BoolArrayX<T> operator>(const T& s) const { BoolArrayX<T> r; r.b.resize(v.size()); for (size_t i = 0; i < v.size(); i++) r.b[i] = v[i] > s; return r; }
  ArrayX inverse() const { ArrayX r; r.v.resize(v.size()); for (size_t i = 0; i < v.size(); i++) r.v[i] = T(1) / v[i]; return r; }
};
template <class T> Matrix<T, Dynamic, 1> BoolArrayX<T>::select(const ArrayX<T>& a, const T& other) const {
  Matrix<T, Dynamic, 1> r((int)b.size());
  for (size_t i = 0; i < b.size(); i++) r[(int)i] = b[i] ? a.v[i] : other;
  return r;
}
template <class D> ArrayX<typename MatrixBase<D>::Scalar> MatrixBase<D>::array() const {
  ArrayX<Scalar> r; r.v.resize(size()); for (int i = 0; i < size(); i++) r.v[i] = lin(i); return r;
}
template <class D> Matrix<typename MatrixBase<D>::Scalar, Dynamic, Dynamic> MatrixBase<D>::asDiagonal() const {
  Matrix<Scalar, Dynamic, Dynamic> r(size(), size()); r.setZero();
  for (int i = 0; i < size(); i++) r.coeffRef(i, i) = lin(i);
  return r;
}

using Vector3d = Matrix<double, 3, 1>;
using Vector4d = Matrix<double, 4, 1>;
using Vector3f = Matrix<float, 3, 1>;
using Matrix3d = Matrix<double, 3, 3>;
using Matrix4d = Matrix<double, 4, 4>;
using Matrix3f = Matrix<float, 3, 3>;
using Matrix4f = Matrix<float, 4, 4>;
using MatrixXd = Matrix<double, Dynamic, Dynamic>;
using VectorXd = Matrix<double, Dynamic, 1>;

// ---- Quaternion (Eigen 3.3 Geometry/Quaternion.h) ------------------------------------------
template <class T> class Quaternion {
  Matrix<T, 4, 1> c_;  // x, y, z, w
 public:
  using Scalar = T;
  Quaternion() {}
  Quaternion(const T& w, const T& x, const T& y, const T& z) { c_[0] = x; c_[1] = y; c_[2] = z; c_[3] = w; }
  Quaternion(const Quaternion&) = default;
  Quaternion& operator=(const Quaternion&) = default;
  // from a rotation matrix: quaternionbase_assign_impl<Other,3,3>
  template <class O> explicit Quaternion(const MatrixBase<O>& mat) { *this = mat; }
  template <class O> Quaternion& operator=(const MatrixBase<O>& a_mat) {
    using std::sqrt;
    auto mat = a_mat.eval();
    T t = mat.trace();
    if (t > T(0)) {
      t = sqrt(t + T(1.0));
      w() = T(0.5) * t;
      t = T(0.5) / t;
      x() = (mat.coeff(2, 1) - mat.coeff(1, 2)) * t;
      y() = (mat.coeff(0, 2) - mat.coeff(2, 0)) * t;
      z() = (mat.coeff(1, 0) - mat.coeff(0, 1)) * t;
    } else {
      int i = 0;
      if (mat.coeff(1, 1) > mat.coeff(0, 0)) i = 1;
      if (mat.coeff(2, 2) > mat.coeff(i, i)) i = 2;
      int j = (i + 1) % 3, k = (j + 1) % 3;
      t = sqrt(mat.coeff(i, i) - mat.coeff(j, j) - mat.coeff(k, k) + T(1.0));
      c_[i] = T(0.5) * t;
      t = T(0.5) / t;
      w() = (mat.coeff(k, j) - mat.coeff(j, k)) * t;
      c_[j] = (mat.coeff(j, i) + mat.coeff(i, j)) * t;
      c_[k] = (mat.coeff(k, i) + mat.coeff(i, k)) * t;
    }
    return *this;
  }
  T x() const { return c_[0]; }  T& x() { return c_[0]; }
  T y() const { return c_[1]; }  T& y() { return c_[1]; }
  T z() const { return c_[2]; }  T& z() { return c_[2]; }
  T w() const { return c_[3]; }  T& w() { return c_[3]; }
  Matrix<T, 4, 1>& coeffs() { return c_; }
  const Matrix<T, 4, 1>& coeffs() const { return c_; }
  Matrix<T, 3, 1> vec() const { return Matrix<T, 3, 1>(c_[0], c_[1], c_[2]); }
  Quaternion& setIdentity() { c_[0] = T(0); c_[1] = T(0); c_[2] = T(0); c_[3] = T(1); return *this; }
  static Quaternion Identity() { Quaternion q; q.setIdentity(); return q; }
  // 4-vector reductions: sequential here (Eigen with SSE2 adds (x²+z²)+(y²+w²); ulp-level, stated)
  T squaredNorm() const { return c_.squaredNorm(); }
  T norm() const { return c_.norm(); }
  void normalize() { c_.normalize(); }
  Quaternion normalized() const { Quaternion q(*this); q.c_ = c_ / c_.norm(); return q; }
  Quaternion conjugate() const { return Quaternion(w(), -x(), -y(), -z()); }
  Quaternion inverse() const {
    T n2 = squaredNorm();
    if (n2 > T(0)) { Quaternion q = conjugate(); q.c_ = q.c_ / n2; return q; }
    Quaternion q; q.c_.setZero(); return q;
  }
  T dot(const Quaternion& o) const { return c_.dot(o.c_); }
  Quaternion operator*(const Quaternion& b) const {
    const Quaternion& a = *this;
    return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                      a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                      a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                      a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
  }
  Quaternion& operator*=(const Quaternion& b) { *this = *this * b; return *this; }
  // _transformVector
  template <class O, class = typename std::enable_if<traits<O>::Cols == 1>::type>
  Matrix<T, 3, 1> operator*(const MatrixBase<O>& v_) const {
    Matrix<T, 3, 1> v = v_.eval();
    Matrix<T, 3, 1> uv = vec().cross(v);
    uv += uv;
    return v + w() * uv + vec().cross(uv);
  }
  // quaternion * 3x3 matrix -> rotation matrix product (RotationBase::operator*)
  Matrix<T, 3, 3> operator*(const Matrix<T, 3, 3>& m) const { return toRotationMatrix() * m; }
  Matrix<T, 3, 3> toRotationMatrix() const {
    Matrix<T, 3, 3> res;
    const T tx = T(2) * x(), ty = T(2) * y(), tz = T(2) * z();
    const T twx = tx * w(), twy = ty * w(), twz = tz * w();
    const T txx = tx * x(), txy = ty * x(), txz = tz * x();
    const T tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
    res.coeffRef(0, 0) = T(1) - (tyy + tzz);
    res.coeffRef(0, 1) = txy - twz;
    res.coeffRef(0, 2) = txz + twy;
    res.coeffRef(1, 0) = txy + twz;
    res.coeffRef(1, 1) = T(1) - (txx + tzz);
    res.coeffRef(1, 2) = tyz - twx;
    res.coeffRef(2, 0) = txz - twy;
    res.coeffRef(2, 1) = tyz + twx;
    res.coeffRef(2, 2) = T(1) - (txx + tyy);
    return res;
  }
  Matrix<T, 3, 3> matrix() const { return toRotationMatrix(); }
  template <class U> Quaternion<U> cast() const { return Quaternion<U>(U(w()), U(x()), U(y()), U(z())); }
  T angularDistance(const Quaternion& other) const {
    using std::atan2; using std::abs;
    Quaternion d = (*this) * other.conjugate();
    return T(2) * atan2(d.vec().norm(), abs(d.w()));
  }
  Quaternion slerp(const T& t, const Quaternion& other) const {
    using std::acos; using std::sin; using std::abs;
    const T one = T(1) - std::numeric_limits<T>::epsilon();
    T d = this->dot(other);
    T absD = abs(d);
    T scale0, scale1;
    if (absD >= one) { scale0 = T(1) - t; scale1 = t; }
    else {
      T theta = acos(absD);
      T sinTheta = sin(theta);
      scale0 = sin((T(1) - t) * theta) / sinTheta;
      scale1 = sin((t * theta)) / sinTheta;
    }
    if (d < T(0)) scale1 = -scale1;
    Quaternion q; q.c_ = scale0 * c_ + scale1 * other.c_;
    return q;
  }
};
using Quaterniond = Quaternion<double>;
using Quaternionf = Quaternion<float>;

// ---- decompositions (restated algorithms) --------------------------------------------------
// Symmetric eigen-decomposition by cyclic Jacobi: eigenvalues ascending, unit eigenvectors in
// columns (sign arbitrary, as in Eigen).
template <class M> class SelfAdjointEigenSolver {
  using S = typename traits<M>::Scalar;
  Matrix<S, traits<M>::Rows, 1> ev_;
  M V_;
 public:
  template <class O> explicit SelfAdjointEigenSolver(const MatrixBase<O>& a_) { compute(a_); }
  template <class O> void compute(const MatrixBase<O>& a_) {
    const int n = a_.rows();
    std::vector<double> A((size_t)n * n), U((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) A[(size_t)i * n + j] = a_.derived().coeff(i, j); U[(size_t)i * n + i] = 1; }
    for (int sweep = 0; sweep < 64; sweep++) {
      double off = 0, diag = 0;
      for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { double v = A[(size_t)i * n + j]; if (i == j) diag += v * v; else if (j > i) off += v * v; }
      if (off == 0.0 || off <= 1e-32 * diag) break;
      for (int p = 0; p < n - 1; p++) for (int q = p + 1; q < n; q++) {
        double apq = A[(size_t)p * n + q];
        if (apq == 0.0) continue;
        double theta = (A[(size_t)q * n + q] - A[(size_t)p * n + p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++) { double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q]; A[(size_t)k * n + p] = c * akp - s * akq; A[(size_t)k * n + q] = s * akp + c * akq; }
        for (int k = 0; k < n; k++) { double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k]; A[(size_t)p * n + k] = c * apk - s * aqk; A[(size_t)q * n + k] = s * apk + c * aqk; }
        for (int k = 0; k < n; k++) { double ukp = U[(size_t)k * n + p], ukq = U[(size_t)k * n + q]; U[(size_t)k * n + p] = c * ukp - s * ukq; U[(size_t)k * n + q] = s * ukp + c * ukq; }
      }
    }
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return A[(size_t)a * n + a] < A[(size_t)b * n + b]; });
    ev_.resize(n, 1); V_.resize(n, n);
    for (int c = 0; c < n; c++) { ev_[c] = A[(size_t)order[c] * n + order[c]]; for (int r = 0; r < n; r++) V_.coeffRef(r, c) = U[(size_t)r * n + order[c]]; }
  }
  const Matrix<S, traits<M>::Rows, 1>& eigenvalues() const { return ev_; }
  const M& eigenvectors() const { return V_; }
};

// Householder QR with column pivoting, least-squares solve (rank from |R_kk| > eps*diagSize*max).
template <class M> class ColPivHouseholderQR {
  int m, n; std::vector<double> A; std::vector<int> perm; std::vector<double> tau; int rank_; std::vector<double> Rd;
 public:
  template <class O> explicit ColPivHouseholderQR(const MatrixBase<O>& a_) : m(a_.rows()), n(a_.cols()), A((size_t)m * n), perm(n), tau(n, 0.0), rank_(0), Rd(n, 0.0) {
    for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) A[(size_t)i * n + j] = a_.derived().coeff(i, j);
    for (int j = 0; j < n; j++) perm[j] = j;
    int steps = std::min(m, n); double maxpivot = 0;
    for (int k = 0; k < steps; k++) {
      int best = k; double bestn = -1;
      for (int j = k; j < n; j++) { double s = 0; for (int i = k; i < m; i++) s += A[(size_t)i * n + j] * A[(size_t)i * n + j]; if (s > bestn) { bestn = s; best = j; } }
      if (best != k) { for (int i = 0; i < m; i++) std::swap(A[(size_t)i * n + k], A[(size_t)i * n + best]); std::swap(perm[k], perm[best]); }
      double tail = 0; for (int i = k + 1; i < m; i++) tail += A[(size_t)i * n + k] * A[(size_t)i * n + k];
      double c0 = A[(size_t)k * n + k], beta;
      if (tail <= 1e-300) { tau[k] = 0; beta = c0; }
      else {
        beta = std::sqrt(c0 * c0 + tail); if (c0 >= 0) beta = -beta;
        for (int i = k + 1; i < m; i++) A[(size_t)i * n + k] /= (c0 - beta);
        tau[k] = (beta - c0) / beta;
      }
      if (tau[k] != 0) for (int j = k + 1; j < n; j++) {
        double s = A[(size_t)k * n + j]; for (int i = k + 1; i < m; i++) s += A[(size_t)i * n + k] * A[(size_t)i * n + j];
        s *= tau[k];
        A[(size_t)k * n + j] -= s; for (int i = k + 1; i < m; i++) A[(size_t)i * n + j] -= s * A[(size_t)i * n + k];
      }
      A[(size_t)k * n + k] = beta; Rd[k] = beta; maxpivot = std::max(maxpivot, std::fabs(beta));
    }
    double thr = std::numeric_limits<double>::epsilon() * steps * maxpivot;
    for (int k = 0; k < steps; k++) if (std::fabs(Rd[k]) > thr) rank_++;
  }
  template <class O> Matrix<double, traits<M>::Cols, 1> solve(const MatrixBase<O>& b_) const {
    std::vector<double> b(m); for (int i = 0; i < m; i++) b[i] = b_.lin(i);
    int steps = std::min(m, n);
    for (int k = 0; k < steps; k++) if (tau[k] != 0) {
      double s = b[k]; for (int i = k + 1; i < m; i++) s += A[(size_t)i * n + k] * b[i];
      s *= tau[k]; b[k] -= s; for (int i = k + 1; i < m; i++) b[i] -= s * A[(size_t)i * n + k];
    }
    std::vector<double> y(n, 0.0);
    for (int k = rank_ - 1; k >= 0; k--) { double s = b[k]; for (int j = k + 1; j < rank_; j++) s -= A[(size_t)k * n + j] * y[j]; y[k] = s / A[(size_t)k * n + k]; }
    Matrix<double, traits<M>::Cols, 1> x; x.resize(n, 1);
    for (int k = 0; k < n; k++) x[perm[k]] = (k < rank_) ? y[k] : 0.0;
    return x;
  }
};

// One-sided Jacobi SVD: A = U S V^T, singular values descending.
template <class M> class JacobiSVD {
  using S = typename traits<M>::Scalar;
  Matrix<S, Dynamic, Dynamic> U_, V_; Matrix<S, Dynamic, 1> sv_;
  unsigned opt_ = 0;
 public:
  JacobiSVD() {}
  JacobiSVD(int, int, unsigned opt = 0) : opt_(opt) {}
  template <class O> explicit JacobiSVD(const MatrixBase<O>& a, unsigned opt = 0) : opt_(opt) { compute(a); }
  template <class O> JacobiSVD& compute(const MatrixBase<O>& a_, unsigned opt) { opt_ = opt; return compute(a_); }
  // U is only formed when asked for (ComputeFullU / ComputeThinU), as in Eigen
  template <class O> JacobiSVD& compute(const MatrixBase<O>& a_) {
    const int m = a_.rows(), n = a_.cols();
    std::vector<double> A((size_t)m * n), V((size_t)n * n, 0.0);
    for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) A[(size_t)i * n + j] = a_.derived().coeff(i, j);
    for (int i = 0; i < n; i++) V[(size_t)i * n + i] = 1;
    for (int sweep = 0; sweep < 64; sweep++) {
      bool rotated = false;
      for (int p = 0; p < n - 1; p++) for (int q = p + 1; q < n; q++) {
        double al = 0, be = 0, ga = 0;
        for (int i = 0; i < m; i++) { al += A[(size_t)i * n + p] * A[(size_t)i * n + p]; be += A[(size_t)i * n + q] * A[(size_t)i * n + q]; ga += A[(size_t)i * n + p] * A[(size_t)i * n + q]; }
        if (ga == 0.0 || std::fabs(ga) <= 1e-16 * std::sqrt(al * be)) continue;
        rotated = true;
        double zeta = (be - al) / (2.0 * ga);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < m; i++) { double ap = A[(size_t)i * n + p], aq = A[(size_t)i * n + q]; A[(size_t)i * n + p] = c * ap - s * aq; A[(size_t)i * n + q] = s * ap + c * aq; }
        for (int i = 0; i < n; i++) { double vp = V[(size_t)i * n + p], vq = V[(size_t)i * n + q]; V[(size_t)i * n + p] = c * vp - s * vq; V[(size_t)i * n + q] = s * vp + c * vq; }
      }
      if (!rotated) break;
    }
    std::vector<double> sv(n); std::vector<int> order(n);
    for (int j = 0; j < n; j++) { double s = 0; for (int i = 0; i < m; i++) s += A[(size_t)i * n + j] * A[(size_t)i * n + j]; sv[j] = std::sqrt(s); order[j] = j; }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return sv[a] > sv[b]; });
    sv_.resize(n, 1); V_.resize(n, n);
    const bool wantU = (opt_ & (ComputeFullU | ComputeThinU)) != 0;
    const int ucols = wantU ? ((opt_ & ComputeFullU) || m <= n ? m : n) : 0;
    U_.resize(wantU ? m : 0, ucols); U_.setZero();
    for (int c = 0; c < n; c++) {
      int o = order[c]; sv_[c] = sv[o];
      for (int r = 0; r < n; r++) V_.coeffRef(r, c) = V[(size_t)r * n + o];
    }
    // left vectors of the numerically non-zero singular values; the rest of U completes an orthonormal
    // basis (Gram-Schmidt over the unit vectors), as a full U does
    const double tiny = (n > 0 ? sv[order[0]] : 0.0) * 1e-13;
    int have = 0;
    for (int c = 0; c < std::min(ucols, n); c++) {
      int o = order[c];
      if (!(sv[o] > tiny)) break;
      for (int r = 0; r < m; r++) U_.coeffRef(r, c) = A[(size_t)r * n + o] / sv[o];
      have = c + 1;
    }
    for (int c = have; c < ucols; c++) {
      for (int e = 0; e < m; e++) {
        std::vector<double> v(m, 0.0); v[e] = 1;
        for (int k = 0; k < c; k++) { double d = 0; for (int r = 0; r < m; r++) d += U_.coeff(r, k) * v[r]; for (int r = 0; r < m; r++) v[r] -= d * U_.coeff(r, k); }
        double nn = 0; for (int r = 0; r < m; r++) nn += v[r] * v[r];
        if (nn > 1e-6) { nn = std::sqrt(nn); for (int r = 0; r < m; r++) U_.coeffRef(r, c) = v[r] / nn; break; }
      }
    }
    return *this;
  }
  const Matrix<S, Dynamic, 1>& singularValues() const { return sv_; }
  const Matrix<S, Dynamic, Dynamic>& matrixU() const { return U_; }
  const Matrix<S, Dynamic, Dynamic>& matrixV() const { return V_; }
};

// Cholesky A = L L^T
template <class M> class LLT {
  M L_;
 public:
  template <class O> explicit LLT(const MatrixBase<O>& a_) {
    const int n = a_.rows(); L_.resize(n, n); L_.setZero();
    for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) {
      double s = a_.derived().coeff(i, j);
      for (int k = 0; k < j; k++) s -= L_.coeff(i, k) * L_.coeff(j, k);
      L_.coeffRef(i, j) = (i == j) ? std::sqrt(s) : s / L_.coeff(j, j);
    }
  }
  const M& matrixL() const { return L_; }
  M matrixU() const { return L_.transpose(); }
};

template <class D> ColPivHouseholderQR<typename MatrixBase<D>::PlainObject> MatrixBase<D>::colPivHouseholderQr() const {
  return ColPivHouseholderQR<PlainObject>(*this);
}

}  // namespace Eigen

#endif
