// CPU check of the host logic of the Laplacian eigen-solver (secedo_b200/csrc/spectral_host.hpp): the subspace
// iteration is run against a plain-loop backend. TEST INFRASTRUCTURE: the product only instantiates it with the
// CUDA backend of spectral.cu.
//   usage: spectral_host_check in.bin out.bin      in: u32 n, u32 k, f64 tol, f64 A[n*n]
//                                                  out: u32 converged, u32 outer, f64 lam[k], f64 Q[n*k] (row-major)
#include "../secedo_b200/csrc/spectral_host.hpp"

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>

struct HostBackend {
    uint32_t ld;
    std::vector<double> M; // ld x ld
    std::vector<std::unique_ptr<double[]>> owned;
    uint64_t mv_cols = 0;
    double *alloc(size_t count) {
        owned.emplace_back(new double[count]());
        return owned.back().get();
    }
    int upload(double *dst, const double *src, size_t count) {
        std::copy(src, src + count, dst);
        return 0;
    }
    int upload_cols(double *dst, int dw, const double *src, int nc, uint32_t rows) {
        for (uint32_t r = 0; r < rows; ++r)
            for (int c = 0; c < nc; ++c) dst[(size_t)r * dw + c] = src[(size_t)r * nc + c];
        return 0;
    }
    int copy_cols(double *dst, int dw, const double *src, int sw, int nc, uint32_t rows) {
        for (uint32_t r = 0; r < rows; ++r)
            for (int c = 0; c < nc; ++c) dst[(size_t)r * dw + c] = src[(size_t)r * sw + c];
        return 0;
    }
    int mv(int width, double *out, const double *in, double alpha, double beta, double gamma, const double *w) {
        std::vector<double> acc((size_t)ld * width, 0.0);
        for (uint32_t k = 0; k < ld; ++k)
            for (uint32_t r = 0; r < ld; ++r) {
                const double m = M[(size_t)k * ld + r];
                if (m == 0.0) continue;
                for (int c = 0; c < width; ++c) acc[(size_t)r * width + c] += m * in[(size_t)k * width + c];
            }
        for (size_t i = 0; i < acc.size(); ++i)
            out[i] = alpha * acc[i] + (beta != 0.0 ? beta * in[i] : 0.0) + (gamma != 0.0 ? gamma * w[i] : 0.0);
        mv_cols += width;
        return 0;
    }
    int gram(const double *X, int p, const double *Y, int q, std::vector<double> &h) {
        h.assign((size_t)p * q, 0.0);
        for (uint32_t r = 0; r < ld; ++r)
            for (int i = 0; i < p; ++i)
                for (int j = 0; j < q; ++j) h[(size_t)i * q + j] += X[(size_t)r * p + i] * Y[(size_t)r * q + j];
        return 0;
    }
    int xr(double *Z, const double *X1, int p1, const std::vector<double> &R1, const double *X2, int p2,
           const std::vector<double> &R2, int q) {
        for (uint32_t r = 0; r < ld; ++r)
            for (int j = 0; j < q; ++j) {
                double s = 0.0;
                for (int k = 0; k < p1; ++k) s += X1[(size_t)r * p1 + k] * R1[(size_t)k * q + j];
                if (X2)
                    for (int k = 0; k < p2; ++k) s += X2[(size_t)r * p2 + k] * R2[(size_t)k * q + j];
                Z[(size_t)r * q + j] = s;
            }
        return 0;
    }
    std::vector<double> lz_alpha, lz_beta;
    int lanczos_tail(const double *v, const double *vp, double *w, int width, int j) {
        double a = 0.0;
        for (uint32_t r = 0; r < ld; ++r) a += v[(size_t)r * width] * w[(size_t)r * width];
        const double bp = j > 0 ? lz_beta[j - 1] : 0.0;
        double nn = 0.0;
        for (uint32_t r = 0; r < ld; ++r) {
            double &x = w[(size_t)r * width];
            x = x - a * v[(size_t)r * width] - bp * vp[(size_t)r * width];
            nn += x * x;
        }
        const double b = std::sqrt(nn);
        for (uint32_t r = 0; r < ld; ++r) w[(size_t)r * width] = b > 1e-14 ? w[(size_t)r * width] / b : 0.0;
        lz_alpha.resize(j + 1);
        lz_beta.resize(j + 1);
        lz_alpha[j] = a;
        lz_beta[j] = b;
        return 0;
    }
    int lanczos_fetch(int steps, std::vector<double> &al, std::vector<double> &be) {
        al.assign(lz_alpha.begin(), lz_alpha.begin() + steps);
        be.assign(lz_beta.begin(), lz_beta.begin() + steps);
        return 0;
    }
    int rank_update(const double *Q, int kq, int off, int nl, const double *lam) {
        for (uint32_t r = 0; r < ld; ++r)
            for (uint32_t c = 0; c < ld; ++c) {
                double s = 0.0;
                for (int i = 0; i < nl; ++i) s += lam[i] * Q[(size_t)r * kq + off + i] * Q[(size_t)c * kq + off + i];
                M[(size_t)r * ld + c] -= s;
            }
        return 0;
    }
};

// --jacobi in.bin out.bin: in = u32 n, f64 A[n*n] (symmetric); out = f64 w[n] ascending, f64 V[n*n] (columns)
// --degree a c top kth mmax: prints the Chebyshev degree
static int small_problems(int argc, char **argv) {
    if (std::string(argv[1]) == "--degree" && argc == 7) {
        std::printf("%d\n", sgpu_spectral::chebyshev_degree(std::atof(argv[2]), std::atof(argv[3]), std::atof(argv[4]),
                                                            std::atof(argv[5]), std::atoi(argv[6])));
        return 0;
    }
    if (std::string(argv[1]) == "--jacobi" && argc == 4) {
        FILE *f = std::fopen(argv[2], "rb");
        uint32_t n = 0;
        if (!f || std::fread(&n, 4, 1, f) != 1) return 2;
        std::vector<double> A((size_t)n * n), w, V;
        if (std::fread(A.data(), 8, A.size(), f) != A.size()) return 2;
        std::fclose(f);
        sgpu_spectral::jacobi_eigh((int)n, A, w, V);
        f = std::fopen(argv[3], "wb");
        std::fwrite(w.data(), 8, n, f);
        std::fwrite(V.data(), 8, V.size(), f);
        std::fclose(f);
        return 0;
    }
    return 2;
}

int main(int argc, char **argv) {
    if (argc >= 2 && argv[1][0] == '-') return small_problems(argc, argv);
    if (argc < 3) return 2;
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    uint32_t n = 0, k = 0;
    double tol = 0;
    if (std::fread(&n, 4, 1, f) != 1 || std::fread(&k, 4, 1, f) != 1 || std::fread(&tol, 8, 1, f) != 1) return 2;
    std::vector<double> A((size_t)n * n);
    if (std::fread(A.data(), 8, A.size(), f) != A.size()) return 2;
    std::fclose(f);
    const uint32_t ld = (n + 31) / 32 * 32;
    const uint32_t want = std::max(2 * k, k + 8);
    const int b = want <= 8 ? 8 : want <= 16 ? 16 : want <= 32 ? 32 : 64;
    const int kq = (int)((k + 7) / 8 * 8);
    HostBackend bk;
    bk.ld = ld;
    bk.M.assign((size_t)ld * ld, 0.0);
    std::vector<double> d(n, 0.0), s(n), v0(n);
    double total = 0;
    for (uint32_t r = 0; r < n; ++r) {
        for (uint32_t c = 0; c < n; ++c) d[r] += A[(size_t)r * n + c];
        total += d[r];
    }
    for (uint32_t r = 0; r < n; ++r) {
        s[r] = d[r] == 0 ? 0 : 1 / std::sqrt(d[r]);
        v0[r] = std::sqrt(d[r] / total);
    }
    for (uint32_t r = 0; r < n; ++r)
        for (uint32_t c = 0; c < n; ++c) bk.M[(size_t)r * ld + c] = (s[r] * s[c]) * A[(size_t)r * n + c] - v0[r] * v0[c];
    double *Q = bk.alloc((size_t)ld * kq);
    for (uint32_t r = 0; r < n; ++r) Q[(size_t)r * kq] = v0[r];
    sgpu_spectral::SolveResult res;
    const int rc = sgpu_spectral::subspace_iteration(bk, n, ld, k, b, kq, tol, Q, std::getenv("SPECTRAL_TRACE") != nullptr, &res);
    if (rc != 0) return 3;
    f = std::fopen(argv[2], "wb");
    const uint32_t conv = res.converged ? 1 : 0;
    std::fwrite(&conv, 4, 1, f);
    std::fwrite(&res.outer, 4, 1, f);
    std::vector<double> lam(k, 0.0);
    for (size_t i = 0; i < res.lam.size() && i < k; ++i) lam[i] = res.lam[i];
    std::fwrite(lam.data(), 8, k, f);
    for (uint32_t r = 0; r < n; ++r) std::fwrite(Q + (size_t)r * kq, 8, k, f);
    std::fclose(f);
    std::fprintf(stderr, "converged %u outer %u mv columns %llu max residual %.3e lo %.4e\n", conv, res.outer,
                 (unsigned long long)bk.mv_cols, res.max_residual, res.lo);
    return 0;
}
