// mini_eigen.h -- TEST INFRASTRUCTURE (oracle/): the small part of Eigen's API that the reference's hot-path sources use,
// so that /root/reference/system/{optimizer,tracker}.cpp, datastructures/imgpyramidrgbd.cpp and utils/LGSX.h compile
// VERBATIM into oracle/_ref/ without Eigen (which is not in this image).  Not Eigen code: plain dense
// column-major matrices evaluated eagerly, written for this purpose.  What matters for parity is the reference's own
// arithmetic (which expression is formed from which operands, in float or promoted to double); a lazily evaluated Eigen
// expression and the eager one below perform the same scalar operations per coefficient for everything used here:
//   R * v + T            row i: (R(i,0) v0 + R(i,1) v1) + R(i,2) v2, then + T(i)   (Eigen's coefficient-based 3x3 product)
//   J * J^T * w          (J(i) J(j)) * w
//   a * X + b * Y + ...  left to right
// LDLT follows Eigen/src/Cholesky/LDLT.h (3.3): in-place lower LDL^T with diagonal pivoting, solve = P^T L^-T D^+ L^-1 P b.
#pragma once
#include <immintrin.h>

#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <ostream>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_ALIGN16 alignas(16)

namespace Eigen {

constexpr int Dynamic = -1;

namespace internal {
inline void *aligned_malloc(size_t n)
{
    void *p = nullptr;
    if (posix_memalign(&p, 32, n ? n : 32) != 0) return nullptr;
    return p;
}
inline void aligned_free(void *p) { free(p); }

template <typename T, int R, int C, bool Dyn = (R == Dynamic || C == Dynamic)>
struct Storage;
template <typename T, int R, int C>
struct Storage<T, R, C, false> {
    T d[R * C];
    int rows() const { return R; }
    int cols() const { return C; }
    void resize(int, int) {}
    T *data() { return d; }
    const T *data() const { return d; }
};
template <typename T, int R, int C>
struct Storage<T, R, C, true> {
    std::vector<T> v;
    int r = R == Dynamic ? 0 : R, c = C == Dynamic ? 0 : C;
    int rows() const { return r; }
    int cols() const { return c; }
    void resize(int rr, int cc) { r = rr; c = cc; v.assign((size_t)rr * cc, T(0)); }
    T *data() { return v.data(); }
    const T *data() const { return v.data(); }
};
}  // namespace internal

template <typename T>
using aligned_allocator = std::allocator<T>;

template <typename T, int R, int C>
class Matrix;

template <typename M>
class LDLT;

// writable view of a fixed-size block of a matrix
template <typename M, int BR, int BC>
struct BlockRef {
    typedef typename M::Scalar T;
    M &m;
    int i0, j0;
    BlockRef &operator=(const Matrix<T, BR, BC> &o)
    {
        for (int j = 0; j < BC; ++j)
            for (int i = 0; i < BR; ++i) m(i0 + i, j0 + j) = o(i, j);
        return *this;
    }
    Matrix<T, BR, BC> eval() const
    {
        Matrix<T, BR, BC> o;
        for (int j = 0; j < BC; ++j)
            for (int i = 0; i < BR; ++i) o(i, j) = m(i0 + i, j0 + j);
        return o;
    }
    operator Matrix<T, BR, BC>() const { return eval(); }
};

// comma initialiser: v << a, b, c;
template <typename M>
struct CommaInit {
    M &m;
    int k;
    CommaInit &operator,(typename M::Scalar s)
    {
        m.data()[k++] = s;      // vectors only (what the reference uses it for)
        return *this;
    }
};

template <typename T, int R, int C>
class Matrix {
public:
    typedef T Scalar;
    internal::Storage<T, R, C> st;

    Matrix() { if (R != Dynamic && C != Dynamic) {} }
    Matrix(int n) { init1(n); }                                   // VectorXf(n)
    Matrix(int r, int c) { st.resize(r, c); }                     // MatrixXf(r, c)
    Matrix(T x, T y, T z) { static_assert(R * C == 3, "3 coefficients"); st.d[0] = x; st.d[1] = y; st.d[2] = z; }
    Matrix(T x, T y, T z, T w) { static_assert(R * C == 4, "4 coefficients"); st.d[0] = x; st.d[1] = y; st.d[2] = z; st.d[3] = w; }
    template <int R2, int C2>
    Matrix(const Matrix<T, R2, C2> &o)                            // conversion between static / dynamic shapes
    {
        st.resize(o.rows(), o.cols());
        assert(rows() == o.rows() && cols() == o.cols());
        for (int i = 0; i < rows() * cols(); ++i) data()[i] = o.data()[i];
    }
    template <int R2, int C2>
    Matrix &operator=(const Matrix<T, R2, C2> &o)
    {
        st.resize(o.rows(), o.cols());
        assert(rows() == o.rows() && cols() == o.cols());
        for (int i = 0; i < rows() * cols(); ++i) data()[i] = o.data()[i];
        return *this;
    }

    int rows() const { return st.rows(); }
    int cols() const { return st.cols(); }
    int size() const { return rows() * cols(); }
    T *data() { return st.data(); }
    const T *data() const { return st.data(); }
    T &operator()(int i, int j) { return data()[(size_t)j * rows() + i]; }
    const T &operator()(int i, int j) const { return data()[(size_t)j * rows() + i]; }
    T &operator()(int i) { return data()[i]; }
    const T &operator()(int i) const { return data()[i]; }
    T &operator[](int i) { return data()[i]; }
    const T &operator[](int i) const { return data()[i]; }

    void setZero() { for (int i = 0; i < size(); ++i) data()[i] = T(0); }
    void setIdentity() { setZero(); for (int i = 0; i < rows() && i < cols(); ++i) (*this)(i, i) = T(1); }
    static Matrix Zero() { Matrix m; m.setZero(); return m; }
    static Matrix Zero(int r, int c) { Matrix m; m.st.resize(r, c); m.setZero(); return m; }
    static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
    Matrix &noalias() { return *this; }

    Matrix<T, C, R> transpose() const
    {
        Matrix<T, C, R> t;
        t.st.resize(cols(), rows());
        for (int i = 0; i < rows(); ++i)
            for (int j = 0; j < cols(); ++j) t(j, i) = (*this)(i, j);
        return t;
    }
    T dot(const Matrix &o) const
    {
        T s = data()[0] * o.data()[0];
        for (int i = 1; i < size(); ++i) s += data()[i] * o.data()[i];
        return s;
    }
    T squaredNorm() const { return dot(*this); }
    T norm() const { return std::sqrt(squaredNorm()); }

    // ---- blocks ------------------------------------------------------------------------------------------------
    template <int BR, int BC>
    BlockRef<Matrix, BR, BC> block(int i, int j) { return BlockRef<Matrix, BR, BC>{*this, i, j}; }
    template <int BR, int BC>
    Matrix<T, BR, BC> block(int i, int j) const
    {
        Matrix<T, BR, BC> o;
        for (int b = 0; b < BC; ++b)
            for (int a = 0; a < BR; ++a) o(a, b) = (*this)(i + a, j + b);
        return o;
    }
    template <int BR, int BC>
    BlockRef<Matrix, BR, BC> topLeftCorner() { return BlockRef<Matrix, BR, BC>{*this, 0, 0}; }
    template <int N>
    BlockRef<Matrix, N, 1> head() { return BlockRef<Matrix, N, 1>{*this, 0, 0}; }
    template <int N>
    Matrix<T, N, 1> head() const
    {
        Matrix<T, N, 1> o;
        for (int i = 0; i < N; ++i) o[i] = data()[i];
        return o;
    }

    // ---- columns -------------------------------------------------------------------------------------------------
    struct ColRef {
        Matrix &m;
        int j;
        template <int R2>
        ColRef &operator=(const Matrix<T, R2, 1> &o)
        {
            assert(o.rows() == m.rows());
            // a column index past the end is undefined behaviour in Eigen (the reference's generateColoredPcl does it when more
            // than a fifth of the pixels are edge points, imgpyramidrgbd.cpp:283,316); here the write is dropped
            if (j >= 0 && j < m.cols())
                for (int i = 0; i < m.rows(); ++i) m(i, j) = o[i];
            return *this;
        }
        operator Matrix<T, R, 1>() const
        {
            Matrix<T, R, 1> o;
            o.st.resize(m.rows(), 1);
            for (int i = 0; i < m.rows(); ++i) o[i] = m(i, j);
            return o;
        }
        template <int N>
        Matrix<T, N, 1> head() const
        {
            Matrix<T, N, 1> o;
            for (int i = 0; i < N; ++i) o[i] = m(i, j);
            return o;
        }
        Matrix<T, 1, R> transpose() const { return Matrix<T, R, 1>(*this).transpose(); }
    };
    ColRef col(int j) { return ColRef{*this, j}; }
    struct ConstCol : Matrix<T, R, 1> {
        template <int N>
        Matrix<T, N, 1> head() const
        {
            Matrix<T, N, 1> o;
            for (int i = 0; i < N; ++i) o[i] = this->data()[i];
            return o;
        }
    };
    ConstCol col(int j) const
    {
        ConstCol o;
        o.st.resize(rows(), 1);
        for (int i = 0; i < rows(); ++i) o[i] = (*this)(i, j);
        return o;
    }
    Matrix<T, R, Dynamic> leftCols(int n) const
    {
        Matrix<T, R, Dynamic> o;
        o.st.resize(rows(), n);
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < rows(); ++i) o(i, j) = (*this)(i, j);
        return o;
    }
    void conservativeResize(int r, int c)
    {
        Matrix t;
        t.st.resize(r, c);
        for (int j = 0; j < c && j < cols(); ++j)
            for (int i = 0; i < r && i < rows(); ++i) t(i, j) = (*this)(i, j);
        *this = t;
    }

    // ---- arithmetic ------------------------------------------------------------------------------------------------
    Matrix operator-() const { Matrix o(*this); for (int i = 0; i < size(); ++i) o.data()[i] = -data()[i]; return o; }
    Matrix &operator+=(const Matrix &o) { for (int i = 0; i < size(); ++i) data()[i] += o.data()[i]; return *this; }
    Matrix &operator-=(const Matrix &o) { for (int i = 0; i < size(); ++i) data()[i] -= o.data()[i]; return *this; }
    Matrix &operator/=(T s) { for (int i = 0; i < size(); ++i) data()[i] /= s; return *this; }
    Matrix &operator*=(T s) { for (int i = 0; i < size(); ++i) data()[i] *= s; return *this; }

    CommaInit<Matrix> operator<<(T s)
    {
        data()[0] = s;
        return CommaInit<Matrix>{*this, 1};
    }

    Matrix inverse() const;          // square, general (Gauss-Jordan with partial pivoting)
    LDLT<Matrix> ldlt() const;

private:
    void init1(int n) { if (R == Dynamic && C == 1) st.resize(n, 1); else if (C == Dynamic && R == 1) st.resize(1, n); }
};

template <typename T, int R, int C>
Matrix<T, R, C> operator*(const Matrix<T, R, C> &a, T s) { Matrix<T, R, C> o(a); o *= s; return o; }
template <typename T, int R, int C>
Matrix<T, R, C> operator*(T s, const Matrix<T, R, C> &a) { Matrix<T, R, C> o(a); for (int i = 0; i < o.size(); ++i) o.data()[i] = s * a.data()[i]; return o; }
// float matrices scaled by an int / double literal (e.g. `1 * v`): the scalar is converted to the matrix type first, as Eigen does
template <typename T, int R, int C>
Matrix<T, R, C> operator*(double s, const Matrix<T, R, C> &a) { return (T)s * a; }
template <typename T, int R, int C>
Matrix<T, R, C> operator*(int s, const Matrix<T, R, C> &a) { return (T)s * a; }
template <typename T, int R, int C>
Matrix<T, R, C> operator/(const Matrix<T, R, C> &a, T s) { Matrix<T, R, C> o(a); o /= s; return o; }

// matrix product, coefficient-based: sum over k in ascending order
template <typename T, int R, int K, int K2, int C>
Matrix<T, R, C> operator*(const Matrix<T, R, K> &a, const Matrix<T, K2, C> &b)
{
    Matrix<T, R, C> o;
    o.st.resize(a.rows(), b.cols());
    const int kk = a.cols();
    for (int j = 0; j < b.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) {
            T s = a(i, 0) * b(0, j);
            for (int k = 1; k < kk; ++k) s += a(i, k) * b(k, j);
            o(i, j) = s;
        }
    return o;
}

template <typename T, int R, int K, typename M, int BR, int BC>
Matrix<T, R, BC> operator*(const Matrix<T, R, K> &a, const BlockRef<M, BR, BC> &b) { return a * b.eval(); }
// also mixed static / dynamic shapes (debug code of the reference: VectorXf - Vector3f); the left operand's shape wins
template <typename T, int R, int C, int R2, int C2>
Matrix<T, R, C> operator-(const Matrix<T, R, C> &a, const Matrix<T, R2, C2> &b)
{
    Matrix<T, R, C> o(a);
    for (int i = 0; i < o.size() && i < b.size(); ++i) o.data()[i] -= b.data()[i];
    return o;
}
template <typename T, int R, int C, int R2, int C2>
Matrix<T, R, C> operator+(const Matrix<T, R, C> &a, const Matrix<T, R2, C2> &b)
{
    Matrix<T, R, C> o(a);
    for (int i = 0; i < o.size() && i < b.size(); ++i) o.data()[i] += b.data()[i];
    return o;
}

template <typename T, int R, int C>
Matrix<T, R, C> Matrix<T, R, C>::inverse() const
{
    const int n = rows();
    std::vector<double> a((size_t)n * 2 * n);
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) { a[(size_t)r * 2 * n + c] = (*this)(r, c); a[(size_t)r * 2 * n + n + c] = r == c ? 1.0 : 0.0; }
    for (int k = 0; k < n; ++k) {
        int piv = k;
        for (int r = k + 1; r < n; ++r)
            if (std::fabs(a[(size_t)r * 2 * n + k]) > std::fabs(a[(size_t)piv * 2 * n + k])) piv = r;
        if (piv != k)
            for (int c = 0; c < 2 * n; ++c) std::swap(a[(size_t)piv * 2 * n + c], a[(size_t)k * 2 * n + c]);
        const double d = 1.0 / a[(size_t)k * 2 * n + k];
        for (int c = 0; c < 2 * n; ++c) a[(size_t)k * 2 * n + c] *= d;
        for (int r = 0; r < n; ++r) {
            if (r == k) continue;
            const double f = a[(size_t)r * 2 * n + k];
            if (f != 0.0)
                for (int c = 0; c < 2 * n; ++c) a[(size_t)r * 2 * n + c] -= f * a[(size_t)k * 2 * n + c];
        }
    }
    Matrix o;
    o.st.resize(n, n);
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) o(r, c) = (T)a[(size_t)r * 2 * n + n + c];
    return o;
}

// Eigen/src/Cholesky/LDLT.h (3.3.x): ldlt_inplace<Lower>::unblocked + LDLT::_solve_impl, restated
template <typename M>
class LDLT {
public:
    typedef typename M::Scalar T;
    M m;
    std::vector<int> tr;
    explicit LDLT(const M &a) : m(a), tr(a.rows())
    {
        const int size = m.rows();
        if (size <= 1) { if (size == 1) tr[0] = 0; return; }
        std::vector<T> temp(size);
        for (int k = 0; k < size; ++k) {
            // largest |diagonal| of the remaining part
            int piv = k;
            T big = std::fabs(m(k, k));
            for (int i = k + 1; i < size; ++i)
                if (std::fabs(m(i, i)) > big) { big = std::fabs(m(i, i)); piv = i; }
            tr[k] = piv;
            if (k != piv) {
                const int s = size - piv - 1;
                for (int c = 0; c < k; ++c) std::swap(m(k, c), m(piv, c));              // row(k).head(k) <-> row(piv).head(k)
                for (int r = 0; r < s; ++r) std::swap(m(piv + 1 + r, k), m(piv + 1 + r, piv));   // col(k).tail(s) <-> col(piv).tail(s)
                std::swap(m(k, k), m(piv, piv));
                for (int i = k + 1; i < piv; ++i) { const T tmp = m(i, k); m(i, k) = m(piv, i); m(piv, i) = tmp; }
            }
            const int rs = size - k - 1;
            if (k > 0) {
                for (int c = 0; c < k; ++c) temp[c] = m(c, c) * m(k, c);            // diag.head(k).asDiagonal() * A10^T
                T s = m(k, 0) * temp[0];
                for (int c = 1; c < k; ++c) s += m(k, c) * temp[c];
                m(k, k) -= s;
                for (int r = 0; r < rs; ++r) {                                       // A21 -= A20 * temp.head(k)
                    T t = m(k + 1 + r, 0) * temp[0];
                    for (int c = 1; c < k; ++c) t += m(k + 1 + r, c) * temp[c];
                    m(k + 1 + r, k) -= t;
                }
            }
            const T akk = m(k, k);
            const bool pivot_ok = std::fabs(akk) > T(0);
            if (k == 0 && !pivot_ok) {
                for (int j = 0; j < size; ++j) tr[j] = j;
                break;
            }
            if (rs > 0 && pivot_ok)
                for (int r = 0; r < rs; ++r) m(k + 1 + r, k) /= akk;
        }
    }
    template <typename B>
    B solve(const B &b) const
    {
        const int size = m.rows();
        B x(b);
        for (int c = 0; c < x.cols(); ++c) {
            for (int k = 0; k < size; ++k) std::swap(x(k, c), x(tr[k], c));                 // dst = P b
            for (int i = 0; i < size; ++i)                                                  // L^-1
                for (int j = 0; j < i; ++j) x(i, c) -= m(i, j) * x(j, c);
            const T tol = T(1) / std::numeric_limits<T>::max();                             // D^+ (pseudo-inverse)
            for (int i = 0; i < size; ++i) {
                if (std::fabs(m(i, i)) > tol) x(i, c) /= m(i, i); else x(i, c) = T(0);
            }
            for (int i = size - 1; i >= 0; --i)                                             // L^-T
                for (int j = i + 1; j < size; ++j) x(i, c) -= m(j, i) * x(j, c);
            for (int k = size - 1; k >= 0; --k) std::swap(x(k, c), x(tr[k], c));            // dst = P^T dst
        }
        return x;
    }
};

template <typename T, int R, int C>
LDLT<Matrix<T, R, C>> Matrix<T, R, C>::ldlt() const { return LDLT<Matrix<T, R, C>>(*this); }

template <typename T, int R, int C>
std::ostream &operator<<(std::ostream &o, const Matrix<T, R, C> &m)
{
    for (int i = 0; i < m.rows(); ++i) {
        for (int j = 0; j < m.cols(); ++j) o << (j ? " " : "") << m(i, j);
        if (i + 1 < m.rows()) o << "\n";
    }
    return o;
}

typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<float, Dynamic, Dynamic> MatrixXf;
typedef Matrix<float, 4, Dynamic> Matrix4Xf;
typedef Matrix<float, Dynamic, 1> VectorXf;
typedef Matrix<double, Dynamic, 1> VectorXd;

}  // namespace Eigen
