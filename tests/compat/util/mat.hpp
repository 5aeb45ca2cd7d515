// Stand-in for the reference's util/mat.hpp: the subset of Mat<T> the hot-path boundary uses
// (owning row-major matrix, move-only). See sequenced_data.hpp in this directory for the purpose.
#pragma once

#include <algorithm>
#include <cassert>
#include <cstdint>
#include <memory>

template <typename T>
class Mat {
    std::unique_ptr<T[]> own_;
    T *el_ = nullptr;
    uint32_t r_ = 0, c_ = 0;

  public:
    Mat() = default;
    Mat(uint32_t r, uint32_t c) : own_(new T[static_cast<size_t>(r) * c]), el_(own_.get()), r_(r), c_(c) {}
    Mat(const Mat &) = delete;
    Mat(Mat &&o) noexcept : own_(std::move(o.own_)), el_(o.el_), r_(o.r_), c_(o.c_) {
        o.el_ = nullptr;
        o.r_ = o.c_ = 0;
    }
    Mat &operator=(Mat &&o) noexcept {
        own_ = std::move(o.own_);
        el_ = o.el_;
        r_ = o.r_;
        c_ = o.c_;
        o.el_ = nullptr;
        o.r_ = o.c_ = 0;
        return *this;
    }
    uint32_t rows() const { return r_; }
    uint32_t cols() const { return c_; }
    bool empty() const { return r_ == 0 && c_ == 0; }
    T &operator()(uint32_t row, uint32_t col) {
        assert(row < r_ && col < c_);
        return el_[static_cast<size_t>(row) * c_ + col];
    }
    const T &operator()(uint32_t row, uint32_t col) const {
        assert(row < r_ && col < c_);
        return el_[static_cast<size_t>(row) * c_ + col];
    }
    T *data() { return el_; }
    const T *data() const { return el_; }
    static Mat fill(uint32_t r, uint32_t c, T v) {
        Mat m(r, c);
        std::fill_n(m.el_, static_cast<size_t>(r) * c, v);
        return m;
    }
    static Mat zeros(uint32_t r, uint32_t c) { return fill(r, c, T(0)); }
};

using Matd = Mat<double>;
