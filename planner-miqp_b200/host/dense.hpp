// dense.hpp -- minimal dense containers for the host driver.
//
// The reference keeps ModelParameters / RawResults in Eigen vectors, matrices and tensors
// (reference src/miqp_planner_data.hpp:46-185).  Eigen is not part of this build image, and
// none of its arithmetic is needed on the host (the arithmetic runs on the GPU), so the host
// driver only needs indexable, resizable storage.  Storage is ROW-major, which is what the
// C ABI (include/miqp_b200.h) consumes without a copy.  The member names follow Eigen so that
// reference-side code (`m(r, c)`, `rows()`, `conservativeResize`, `setConstant`, ...) keeps
// compiling when a maintainer swaps these aliases for the Eigen types.
#pragma once
#include <algorithm>
#include <array>
#include <cassert>
#include <cstddef>
#include <vector>

namespace miqp {
namespace dense {

template <class T>
class Vec {
 public:
  Vec() = default;
  explicit Vec(int n) : v_(n, T(0)) {}
  int size() const { return (int)v_.size(); }
  int rows() const { return size(); }
  void resize(int n) { v_.assign(n, T(0)); }
  void conservativeResize(int n) { v_.resize(n, T(0)); }
  void setZero() { std::fill(v_.begin(), v_.end(), T(0)); }
  void setConstant(T c) { std::fill(v_.begin(), v_.end(), c); }
  T &operator()(int i) { assert(i >= 0 && i < size()); return v_[i]; }
  const T &operator()(int i) const { assert(i >= 0 && i < size()); return v_[i]; }
  T &operator[](int i) { return v_[i]; }
  const T &operator[](int i) const { return v_[i]; }
  T *data() { return v_.data(); }
  const T *data() const { return v_.data(); }
  T maxCoeff() const { return *std::max_element(v_.begin(), v_.end()); }
  T minCoeff() const { return *std::min_element(v_.begin(), v_.end()); }
  bool operator==(const Vec &o) const { return v_ == o.v_; }

 private:
  std::vector<T> v_;
};

template <class T>
class Mat {
 public:
  Mat() = default;
  Mat(int r, int c) : r_(r), c_(c), v_((size_t)r * c, T(0)) {}
  int rows() const { return r_; }
  int cols() const { return c_; }
  int size() const { return r_ * c_; }
  void resize(int r, int c) { r_ = r; c_ = c; v_.assign((size_t)r * c, T(0)); }
  // keeps the overlapping top-left block, new entries are zero
  void conservativeResize(int r, int c) {
    if (r == r_ && c == c_) return;
    std::vector<T> n((size_t)r * c, T(0));
    for (int i = 0; i < std::min(r, r_); ++i)
      for (int j = 0; j < std::min(c, c_); ++j) n[(size_t)i * c + j] = v_[(size_t)i * c_ + j];
    v_.swap(n); r_ = r; c_ = c;
  }
  void setZero() { std::fill(v_.begin(), v_.end(), T(0)); }
  void setConstant(T c) { std::fill(v_.begin(), v_.end(), c); }
  T &operator()(int i, int j) { assert(i >= 0 && i < r_ && j >= 0 && j < c_); return v_[(size_t)i * c_ + j]; }
  const T &operator()(int i, int j) const { assert(i >= 0 && i < r_ && j >= 0 && j < c_); return v_[(size_t)i * c_ + j]; }
  T *data() { return v_.data(); }
  const T *data() const { return v_.data(); }
  T *rowPtr(int i) { return v_.data() + (size_t)i * c_; }
  const T *rowPtr(int i) const { return v_.data() + (size_t)i * c_; }
  template <class It> void setRow(int i, It first) { for (int j = 0; j < c_; ++j, ++first) (*this)(i, j) = *first; }
  T maxCoeff() const { return *std::max_element(v_.begin(), v_.end()); }
  T minCoeff() const { return *std::min_element(v_.begin(), v_.end()); }
  bool operator==(const Mat &o) const { return r_ == o.r_ && c_ == o.c_ && v_ == o.v_; }

 private:
  int r_ = 0, c_ = 0;
  std::vector<T> v_;
};

// rank-R tensor, row-major (last index fastest)
template <class T, int R>
class Tensor {
 public:
  Tensor() { d_.fill(0); }
  template <class... I> void resize(I... dims) {
    static_assert(sizeof...(I) == R, "rank mismatch");
    d_ = {{(int)dims...}};
    size_t n = 1; for (int k = 0; k < R; ++k) n *= (size_t)d_[k];
    v_.assign(n, T(0));
  }
  int dimension(int k) const { return d_[k]; }
  size_t size() const { return v_.size(); }
  void setZero() { std::fill(v_.begin(), v_.end(), T(0)); }
  void setConstant(T c) { std::fill(v_.begin(), v_.end(), c); }
  template <class... I> T &operator()(I... idx) { return v_[offset(idx...)]; }
  template <class... I> const T &operator()(I... idx) const { return v_[offset(idx...)]; }
  T *data() { return v_.data(); }
  const T *data() const { return v_.data(); }

 private:
  template <class... I> size_t offset(I... idx) const {
    static_assert(sizeof...(I) == R, "rank mismatch");
    const int ix[R] = {(int)idx...};
    size_t o = 0;
    for (int k = 0; k < R; ++k) { assert(ix[k] >= 0 && ix[k] < d_[k]); o = o * (size_t)d_[k] + (size_t)ix[k]; }
    return o;
  }
  std::array<int, R> d_;
  std::vector<T> v_;
};

using VectorXd = Vec<double>;
using VectorXi = Vec<int>;
using MatrixXd = Mat<double>;
using MatrixXi = Mat<int>;

}  // namespace dense
}  // namespace miqp
