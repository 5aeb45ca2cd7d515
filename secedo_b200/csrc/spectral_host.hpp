// Host-side pieces of the Laplacian eigen-solver (spectral.cu): the small dense problems of the
// Rayleigh-Ritz steps and the choice of the Chebyshev filter. Plain C++ (no CUDA), so that the CPU test
// suite can compile and check it without a GPU (tests/test_spectral_host.py).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

namespace sgpu_spectral {

// Cyclic Jacobi for a symmetric n x n matrix (row-major, overwritten). Eigenvalues ascending in w, the
// matching eigenvectors in the COLUMNS of V (row-major n x n). n is at most a few dozen here.
inline void jacobi_eigh(int n, std::vector<double> &A, std::vector<double> &w, std::vector<double> &V) {
    V.assign(static_cast<size_t>(n) * n, 0.0);
    for (int i = 0; i < n; ++i) {
        V[static_cast<size_t>(i) * n + i] = 1.0;
    }
    auto a = [&](int i, int j) -> double & { return A[static_cast<size_t>(i) * n + j]; };
    auto v = [&](int i, int j) -> double & { return V[static_cast<size_t>(i) * n + j]; };
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; ++i) {
            diag += a(i, i) * a(i, i);
            for (int j = i + 1; j < n; ++j) {
                off += a(i, j) * a(i, j);
            }
        }
        if (off <= 1e-32 * diag || off == 0.0) {
            break;
        }
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                const double apq = a(p, q);
                if (apq == 0.0) {
                    continue;
                }
                const double app = a(p, p), aqq = a(q, q);
                // negligible rotation: off-diagonal element far below both diagonal elements
                if (std::fabs(apq) < 1e-300) {
                    a(p, q) = a(q, p) = 0.0;
                    continue;
                }
                const double tau = (aqq - app) / (2.0 * apq);
                const double t = (tau >= 0.0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
                for (int k = 0; k < n; ++k) { // columns p, q of A
                    const double akp = a(k, p), akq = a(k, q);
                    a(k, p) = c * akp - s * akq;
                    a(k, q) = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) { // rows p, q of A
                    const double apk = a(p, k), aqk = a(q, k);
                    a(p, k) = c * apk - s * aqk;
                    a(q, k) = s * apk + c * aqk;
                }
                a(p, q) = a(q, p) = 0.0;
                for (int k = 0; k < n; ++k) {
                    const double vkp = v(k, p), vkq = v(k, q);
                    v(k, p) = c * vkp - s * vkq;
                    v(k, q) = s * vkp + c * vkq;
                }
            }
        }
    }
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) {
        order[i] = i;
    }
    std::sort(order.begin(), order.end(), [&](int x, int y) { return a(x, x) < a(y, y); });
    w.resize(n);
    std::vector<double> Vs(static_cast<size_t>(n) * n);
    for (int j = 0; j < n; ++j) {
        w[j] = a(order[j], order[j]);
        for (int i = 0; i < n; ++i) {
            Vs[static_cast<size_t>(i) * n + j] = v(i, order[j]);
        }
    }
    V.swap(Vs);
}

// Degree of the Chebyshev filter that damps [a, c] for one outer iteration.
//   theta_top: largest Ritz value not yet locked; theta_k: smallest Ritz value still wanted (both > c).
// The polynomial grows like cosh(m acosh(t)) outside the interval, t = (theta - centre) / half width.
// Two limits: (1) the growth at theta_top relative to theta_k stays below 2e6, so that the directions
// that converge first cannot swamp the others inside one iteration (they are locked and deflated before
// the degree rises); (2) no more growth at theta_k than a full convergence needs (2e12).
inline int chebyshev_degree(double a, double c, double theta_top, double theta_k, int m_max) {
    const double e = 0.5 * (c - a), ctr = 0.5 * (c + a);
    if (!(e > 0.0)) {
        return 2;
    }
    const double t_top = std::max((theta_top - ctr) / e, 1.0 + 1e-12);
    const double t_k = std::max((theta_k - ctr) / e, 1.0 + 1e-12);
    const double da = std::acosh(t_top) - std::acosh(t_k);
    double m = m_max;
    if (da > 0.0) {
        m = std::min(m, std::log(2e6) / da);
    }
    m = std::min(m, std::ceil(std::log(2e12) / std::acosh(t_k)));
    return std::max(2, static_cast<int>(m));
}

// Bounds of the spectrum from k Lanczos steps (alpha[0..k), beta[0..k): beta[j] = norm of the j-th
// residual): extreme Ritz values moved outwards by their residual bound |beta_k y_k| plus 1 % of the width.
inline void lanczos_bounds(const std::vector<double> &alpha, const std::vector<double> &beta, double *lo, double *hi) {
    const int k = static_cast<int>(alpha.size());
    std::vector<double> T(static_cast<size_t>(k) * k, 0.0), w, Y;
    for (int i = 0; i < k; ++i) {
        T[static_cast<size_t>(i) * k + i] = alpha[i];
        if (i + 1 < k) {
            T[static_cast<size_t>(i) * k + i + 1] = T[static_cast<size_t>(i + 1) * k + i] = beta[i];
        }
    }
    jacobi_eigh(k, T, w, Y);
    const double bk = std::fabs(beta[k - 1]);
    const double width = w[k - 1] - w[0];
    *lo = w[0] - bk * std::fabs(Y[static_cast<size_t>(k - 1) * k + 0]) - 0.01 * width;
    *hi = w[k - 1] + bk * std::fabs(Y[static_cast<size_t>(k - 1) * k + (k - 1)]) + 0.01 * width;
}

// splitmix64 -> uniform in (-1, 1); the start block only has to be generic
struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    double next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        return (static_cast<double>(z >> 11) + 0.5) * (2.0 / 9007199254740992.0) - 1.0;
    }
};

// ------------------------------------------------------------------------------------------------------------
// The subspace iteration itself, written against a backend that owns the O(N^2) and O(N b^2) work (spectral.cu:
// CUDA kernels; tests/spectral_host_check.cpp: plain loops, to check this logic without a GPU). Blocks are
// row-major [ld][width] arrays of the backend; the operator M = B - v0 v0^T is held by the backend.
//
//   bk.alloc(count) -> double*                       bk.upload(dst, src, count)
//   bk.upload_cols(dst, dst_width, src, ncols, rows)  host [rows][ncols] into the first ncols columns of dst
//   bk.copy_cols(dst, dst_width, src, src_width, ncols, rows)
//   bk.mv(width, out, in, alpha, beta, gamma, w)      out = alpha M in + beta in + gamma w
//   bk.gram(X, p, Y, q, host)                         host = X^T Y (p x q)
//   bk.xr(Z, X1, p1, R1, X2, p2, R2, q)               Z = X1 R1 + X2 R2 (X2 may be null)
//   bk.lanczos_tail(v, v_prev, w, width, j)           one Lanczos step after w = M v, on column 0 (see below)
//   bk.lanczos_fetch(steps, alpha, beta)              the recurrence coefficients of all steps
//   bk.rank_update(Q, kq, off, nl, coef)              M -= sum_i coef[i] q_(off+i) q_(off+i)^T
// Every call returns 0 or an error code that is passed up unchanged.
struct SolveResult {
    std::vector<double> lam; // eigenvalues of B of the locked columns of Q (lam[0] = 1 for v0)
    uint32_t outer = 0;
    double lo = -1.0, hi = 1.0, max_residual = 0.0, last_residual = 0.0;
    bool converged = false;
};

inline std::vector<double> scaled_identity(int p, double s = 1.0) {
    std::vector<double> r(static_cast<size_t>(p) * p, 0.0);
    for (int i = 0; i < p; ++i) {
        r[static_cast<size_t>(i) * p + i] = s;
    }
    return r;
}

#define SGPU_SP_TRY(expr)    \
    do {                     \
        int rc__ = (expr);   \
        if (rc__ != 0) {     \
            return rc__;     \
        }                    \
    } while (0)

// Q: [ld][kq], column 0 = v0 (unit norm), other columns zero. On return columns 0..k-1 hold the eigenvectors
// in the order of res->lam. n = true dimension (rows >= n of every block stay zero), b = block width.
template <class BK>
int subspace_iteration(BK &bk, uint32_t n, uint32_t ld, uint32_t k, int b, int kq, double tol, double *Q, bool trace,
                       SolveResult *res) {
    res->lam.assign(1, 1.0);
    res->converged = k <= 1;
    if (k <= 1) {
        return 0;
    }
    Rng rng(0x5ECED0ull + n);
    const std::vector<double> none;
    // ---- lower end of the spectrum of M: Lanczos on a block of width 8 whose column 0 carries the vector ----
    {
        const int lw = 8, steps = static_cast<int>(std::min<uint32_t>(24, n - 1));
        const size_t lb = static_cast<size_t>(ld) * lw;
        double *V = bk.alloc(lb), *Vp = bk.alloc(lb), *Wv = bk.alloc(lb);
        if (!V || !Vp || !Wv) {
            return -1;
        }
        std::vector<double> hv(lb, 0.0), gm;
        for (uint32_t r = 0; r < n; ++r) {
            hv[static_cast<size_t>(r) * lw] = rng.next();
        }
        SGPU_SP_TRY(bk.upload(V, hv.data(), lb));
        std::fill(hv.begin(), hv.end(), 0.0);
        SGPU_SP_TRY(bk.upload(Vp, hv.data(), lb));
        SGPU_SP_TRY(bk.gram(V, lw, V, lw, gm));
        SGPU_SP_TRY(bk.xr(Wv, V, lw, scaled_identity(lw, 1.0 / std::sqrt(gm[0])), nullptr, 0, none, lw));
        std::swap(V, Wv);
        // three-term recurrence without host round trips: the backend keeps alpha_j, beta_j
        for (int j = 0; j < steps; ++j) {
            SGPU_SP_TRY(bk.mv(lw, Wv, V, 1.0, 0.0, 0.0, nullptr)); // w = M v
            // alpha_j = v.w; w -= alpha_j v + beta_(j-1) v_prev; beta_j = |w|; w /= beta_j   (column 0)
            SGPU_SP_TRY(bk.lanczos_tail(V, Vp, Wv, lw, j));
            double *old_v = V; // rotate: v_prev <- v, v <- w, w <- scratch
            V = Wv;
            Wv = Vp;
            Vp = old_v;
        }
        std::vector<double> al, be;
        SGPU_SP_TRY(bk.lanczos_fetch(steps, al, be));
        for (size_t j = 0; j < be.size(); ++j) { // breakdown: an invariant subspace was found
            if (!(be[j] > 1e-14)) {
                al.resize(j + 1);
                be.resize(j + 1);
                break;
            }
        }
        lanczos_bounds(al, be, &res->lo, &res->hi);
    }
    const double lo = res->lo, hi = res->hi;
    // Locked directions are moved to the LOWER END of the damped interval (eigenvalue lo, where the filter
    // polynomial has modulus 1), not to 0: when the matrix is nearly constant all other eigenvalues of B are
    // negative (they sum to -1), 0 would lie above the damped interval and rounding noise along v0 would be
    // amplified by the filter until it swamps the block.
    {
        const double shift = 0.0 - lo; // v0 sits at 0 in M = B - v0 v0^T
        SGPU_SP_TRY(bk.rank_update(Q, kq, 0, 1, &shift));
    }
    const size_t blk = static_cast<size_t>(ld) * b;
    double *X = bk.alloc(blk), *Y = bk.alloc(blk), *W = bk.alloc(blk), *Z = bk.alloc(blk);
    if (!X || !Y || !W || !Z) {
        return -1;
    }
    {
        std::vector<double> hx(blk, 0.0);
        for (uint32_t r = 0; r < n; ++r) {
            for (int c = 0; c < b; ++c) {
                hx[static_cast<size_t>(r) * b + c] = rng.next();
            }
        }
        SGPU_SP_TRY(bk.upload(X, hx.data(), blk));
    }
    std::vector<double> C, G, w, U, R1, R2, th(b), rn(b), dn(b);
    const int m_max = 100;
    uint32_t nlock = 1;
    for (;;) {
        ++res->outer;
        // -- orthonormalise the block against the locked vectors and itself (twice) --
        for (int pass = 0; pass < 2; ++pass) {
            SGPU_SP_TRY(bk.gram(Q, kq, X, b, C)); // kq x b
            for (auto &v : C) {
                v = -v;
            }
            SGPU_SP_TRY(bk.xr(Z, X, b, scaled_identity(b), Q, kq, C, b));
            std::swap(X, Z);
            SGPU_SP_TRY(bk.gram(X, b, X, b, G));
            for (int i = 0; i < b; ++i) {
                const double d = G[static_cast<size_t>(i) * b + i];
                dn[i] = d > 0 ? 1.0 / std::sqrt(d) : 0.0;
            }
            for (int i = 0; i < b; ++i) {
                for (int j = 0; j < b; ++j) {
                    G[static_cast<size_t>(i) * b + j] *= dn[i] * dn[j];
                }
            }
            jacobi_eigh(b, G, w, U);
            R1.assign(static_cast<size_t>(b) * b, 0.0);
            for (int i = 0; i < b; ++i) {
                for (int j = 0; j < b; ++j) {
                    R1[static_cast<size_t>(i) * b + j] = dn[i] * U[static_cast<size_t>(i) * b + j] / std::sqrt(std::max(w[j], 1e-300));
                }
            }
            SGPU_SP_TRY(bk.xr(Z, X, b, R1, nullptr, 0, none, b));
            std::swap(X, Z);
        }
        // -- Rayleigh-Ritz --
        SGPU_SP_TRY(bk.mv(b, Y, X, 1.0, 0.0, 0.0, nullptr)); // Y = M X
        SGPU_SP_TRY(bk.gram(X, b, Y, b, G));
        for (int i = 0; i < b; ++i) {
            for (int j = i + 1; j < b; ++j) {
                const double s = 0.5 * (G[static_cast<size_t>(i) * b + j] + G[static_cast<size_t>(j) * b + i]);
                G[static_cast<size_t>(i) * b + j] = G[static_cast<size_t>(j) * b + i] = s;
            }
        }
        jacobi_eigh(b, G, w, U); // ascending
        R1.assign(static_cast<size_t>(b) * b, 0.0);
        R2.assign(static_cast<size_t>(b) * b, 0.0);
        for (int j = 0; j < b; ++j) { // descending order of the Ritz values
            th[j] = w[b - 1 - j];
            for (int i = 0; i < b; ++i) {
                R1[static_cast<size_t>(i) * b + j] = U[static_cast<size_t>(i) * b + (b - 1 - j)];
                R2[static_cast<size_t>(i) * b + j] = -th[j] * R1[static_cast<size_t>(i) * b + j];
            }
        }
        SGPU_SP_TRY(bk.xr(W, Y, b, R1, X, b, R2, b));        // residuals M x - theta x
        SGPU_SP_TRY(bk.xr(Z, X, b, R1, nullptr, 0, none, b)); // Ritz vectors
        std::swap(X, Z);
        SGPU_SP_TRY(bk.gram(W, b, W, b, G));
        for (int j = 0; j < b; ++j) {
            rn[j] = std::sqrt(std::max(G[static_cast<size_t>(j) * b + j], 0.0));
        }
        // -- lock the converged Ritz pairs at the top --
        int nl = 0;
        while (nl < b && nlock + nl < k && rn[nl] <= tol) {
            res->max_residual = std::max(res->max_residual, rn[nl]);
            ++nl;
        }
        if (trace) {
            std::fprintf(stderr, "[spectral] outer %u locked %u+%d theta %.6e %.6e .. %.6e res %.2e %.2e lo %.4e\n", res->outer,
                         nlock, nl, th[0], th[1], th[b - 1], rn[0], rn[1], lo);
        }
        if (nl > 0) {
            SGPU_SP_TRY(bk.copy_cols(Q + nlock, kq, X, b, nl, ld));
            std::vector<double> shift(nl);
            for (int i = 0; i < nl; ++i) {
                shift[i] = th[i] - lo;
            }
            SGPU_SP_TRY(bk.rank_update(Q, kq, static_cast<int>(nlock), nl, shift.data()));
            for (int i = 0; i < nl; ++i) {
                res->lam.push_back(th[i]);
            }
            nlock += nl;
            if (nlock >= k) {
                res->converged = true;
                return 0;
            }
            // fresh random columns in place of the locked ones
            std::vector<double> fresh(static_cast<size_t>(ld) * nl, 0.0);
            for (uint32_t r = 0; r < n; ++r) {
                for (int c = 0; c < nl; ++c) {
                    fresh[static_cast<size_t>(r) * nl + c] = rng.next();
                }
            }
            SGPU_SP_TRY(bk.upload_cols(X, b, fresh.data(), nl, ld));
        }
        res->last_residual = rn[std::min(nl, b - 1)];
        if (res->outer >= 300) {
            return 0; // not converged
        }
        // -- Chebyshev filter on the unlocked part of the spectrum --
        const int nw = static_cast<int>(k - nlock); // still wanted
        const double top = th[nl];
        const double kth = th[std::min(b - 1, nl + nw - 1)];
        double c = th[b - 1];
        if (!(c > lo)) {
            c = lo + 1e-3 * (hi - lo);
        }
        const double e = 0.5 * (c - lo), ctr = 0.5 * (c + lo);
        const int m = chebyshev_degree(lo, c, top, kth, m_max);
        double sig = top > c ? e / (top - ctr) : 0.5;
        const double tau = 2.0 / sig;
        SGPU_SP_TRY(bk.mv(b, Y, X, sig / e, -ctr * sig / e, 0.0, nullptr)); // Y = (M X - ctr X) sig / e
        for (int i = 2; i <= m; ++i) {
            const double sn = 1.0 / (tau - sig);
            // W = (M Y - ctr Y) 2 sn / e - sig sn X
            SGPU_SP_TRY(bk.mv(b, W, Y, 2.0 * sn / e, -ctr * 2.0 * sn / e, -sig * sn, X));
            double *ox = X;
            X = Y;
            Y = W;
            W = ox;
            sig = sn;
        }
        std::swap(X, Y);
    }
}

} // namespace sgpu_spectral
