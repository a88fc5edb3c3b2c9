// eigen_shim.hpp -- the small subset of Eigen's dense API that copra's public interface exposes
// (MatrixXd / VectorXd, column-major, data() / rows() / cols(), comma initialiser, Zero / Ones /
// Identity / Constant, head / tail / segment, maxCoeff ...).  It is used ONLY when <Eigen/Core> is not
// installed (this build image has no Eigen); with real Eigen the facade compiles against it unchanged.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <ostream>
#include <vector>

namespace Eigen {

typedef std::ptrdiff_t Index;

template <class Derived> class CommaInit {
public:
    CommaInit(Derived& m, double first) : m_(m), k_(0) { put(first); }
    CommaInit& operator,(double v) { put(v); return *this; }

private:
    void put(double v)
    {
        const Index r = k_ / m_.cols(), c = k_ % m_.cols(); // Eigen fills row by row
        m_(r, c) = v;
        ++k_;
    }
    Derived& m_;
    Index k_;
};

class MatrixXd {
public:
    MatrixXd() : r_(0), c_(0) {}
    MatrixXd(Index r, Index c) : r_(r), c_(c), a_(size_t(r * c), 0.0) {}
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Index size() const { return r_ * c_; }
    double* data() { return a_.data(); }
    const double* data() const { return a_.data(); }
    double& operator()(Index i, Index j) { return a_[size_t(j * r_ + i)]; }
    double operator()(Index i, Index j) const { return a_[size_t(j * r_ + i)]; }
    void resize(Index r, Index c) { r_ = r; c_ = c; a_.assign(size_t(r * c), 0.0); }
    MatrixXd& setZero() { std::fill(a_.begin(), a_.end(), 0.0); return *this; }
    MatrixXd& setConstant(double v) { std::fill(a_.begin(), a_.end(), v); return *this; }
    MatrixXd& setIdentity()
    {
        setZero();
        for (Index i = 0; i < std::min(r_, c_); ++i) (*this)(i, i) = 1.0;
        return *this;
    }
    static MatrixXd Zero(Index r, Index c) { return MatrixXd(r, c); }
    static MatrixXd Constant(Index r, Index c, double v) { MatrixXd m(r, c); m.setConstant(v); return m; }
    static MatrixXd Ones(Index r, Index c) { return Constant(r, c, 1.0); }
    static MatrixXd Identity(Index r, Index c) { MatrixXd m(r, c); m.setIdentity(); return m; }
    CommaInit<MatrixXd> operator<<(double v) { return CommaInit<MatrixXd>(*this, v); }
    MatrixXd block(Index i, Index j, Index r, Index c) const
    {
        MatrixXd b(r, c);
        for (Index cc = 0; cc < c; ++cc)
            for (Index rr = 0; rr < r; ++rr) b(rr, cc) = (*this)(i + rr, j + cc);
        return b;
    }
    MatrixXd transpose() const
    {
        MatrixXd t(c_, r_);
        for (Index j = 0; j < c_; ++j)
            for (Index i = 0; i < r_; ++i) t(j, i) = (*this)(i, j);
        return t;
    }
    double maxCoeff() const { return *std::max_element(a_.begin(), a_.end()); }
    double minCoeff() const { return *std::min_element(a_.begin(), a_.end()); }
    bool isApprox(const MatrixXd& o, double prec = 1e-12) const
    {
        if (r_ != o.r_ || c_ != o.c_) return false;
        double d = 0, n1 = 0, n2 = 0;
        for (size_t k = 0; k < a_.size(); ++k) { d += (a_[k] - o.a_[k]) * (a_[k] - o.a_[k]); n1 += a_[k] * a_[k]; n2 += o.a_[k] * o.a_[k]; }
        return d <= prec * prec * std::min(n1, n2);
    }
    MatrixXd& operator*=(double s) { for (double& v : a_) v *= s; return *this; }

protected:
    Index r_, c_;
    std::vector<double> a_;
};

class VectorXd {
public:
    VectorXd() {}
    explicit VectorXd(Index n) : a_(size_t(n), 0.0) {}
    Index rows() const { return Index(a_.size()); }
    Index cols() const { return 1; }
    Index size() const { return Index(a_.size()); }
    double* data() { return a_.data(); }
    const double* data() const { return a_.data(); }
    double& operator()(Index i) { return a_[size_t(i)]; }
    double operator()(Index i) const { return a_[size_t(i)]; }
    double& operator()(Index i, Index) { return a_[size_t(i)]; } // for the comma initialiser
    double& operator[](Index i) { return a_[size_t(i)]; }
    double operator[](Index i) const { return a_[size_t(i)]; }
    void resize(Index n) { a_.assign(size_t(n), 0.0); }
    void conservativeResize(Index n) { a_.resize(size_t(n), 0.0); }
    VectorXd& setZero() { std::fill(a_.begin(), a_.end(), 0.0); return *this; }
    VectorXd& setConstant(double v) { std::fill(a_.begin(), a_.end(), v); return *this; }
    VectorXd& setConstant(Index n, double v) { a_.assign(size_t(n), v); return *this; }
    static VectorXd Zero(Index n) { return VectorXd(n); }
    static VectorXd Constant(Index n, double v) { VectorXd x(n); x.setConstant(v); return x; }
    static VectorXd Ones(Index n) { return Constant(n, 1.0); }
    CommaInit<VectorXd> operator<<(double v) { return CommaInit<VectorXd>(*this, v); }
    VectorXd segment(Index i, Index n) const { VectorXd s(n); std::copy(a_.begin() + i, a_.begin() + i + n, s.a_.begin()); return s; }
    VectorXd head(Index n) const { return segment(0, n); }
    VectorXd tail(Index n) const { return segment(size() - n, n); }
    double maxCoeff() const { return *std::max_element(a_.begin(), a_.end()); }
    double minCoeff() const { return *std::min_element(a_.begin(), a_.end()); }
    bool isApprox(const VectorXd& o, double prec = 1e-12) const
    {
        if (a_.size() != o.a_.size()) return false;
        double d = 0, n1 = 0, n2 = 0;
        for (size_t k = 0; k < a_.size(); ++k) { d += (a_[k] - o.a_[k]) * (a_[k] - o.a_[k]); n1 += a_[k] * a_[k]; n2 += o.a_[k] * o.a_[k]; }
        return d <= prec * prec * std::min(n1, n2);
    }

private:
    std::vector<double> a_;
};

inline std::ostream& operator<<(std::ostream& os, const VectorXd& v)
{
    for (Index i = 0; i < v.size(); ++i) os << v(i) << (i + 1 < v.size() ? "\n" : "");
    return os;
}

} // namespace Eigen
