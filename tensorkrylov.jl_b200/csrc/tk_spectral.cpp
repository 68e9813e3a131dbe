// tk_spectral.cpp -- spectral data of the leading minors of A_1 inside the library (host side).
//
// The reference recomputes, in every iteration k, the extreme eigenvalues of the k x k leading minor of the FIRST
// coefficient matrix (SpectralData / update_data! / extreme_eigvals, src/eigenvalues.jl:268-370):
//   SymInstance + Laplace    analytic                                   (:247-265, :335)   -> tk_tables.cpp
//   SymInstance + RandSPD    extrema of eigvals(A_1[1:k,1:k]) times d   (:337)
//   SymInstance + EigValMat  extrema of diag(A_1)[1:k] times d          (:339)
//   NonSymInstance, any      minimum(eigvals(A_1[1:k,1:k])) times d     (:344-350)
// They depend only on A_1, never on the Krylov state, so all of k = 2..nmax is done before the loop: one task per
// minor on the host threads.  Symmetric minors: Householder tridiagonalisation + Sturm bisection for the two
// extremes (tridiagonal minors skip the reduction).  General minors: Householder reduction to Hessenberg form
// (skipped when the minor already is Hessenberg, as for the convection-diffusion operator) + the shifted QR
// iteration on the Hessenberg matrix, eigenvalues only.  Results are cached per process, keyed by the content of the
// leading block: a second handle over the same operator pays nothing.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/tensorkrylov_b200.h"
#include "tk_host.h"

namespace tk {

// ---- symmetric: Householder tridiagonalisation of the lower triangle (a is n x n column-major, destroyed) ----------
static void tridiagonalize(std::vector<double>& a, int n, std::vector<double>& d, std::vector<double>& e) {
    auto A = [&](int i, int j) -> double& { return a[(size_t)j * n + i]; };   // i >= j: lower triangle
    d.assign(n, 0.0);
    e.assign(n, 0.0);                                   // e[i] couples rows i-1 and i
    for (int i = n - 1; i >= 1; --i) {
        const int l = i - 1;
        double h = 0.0, scale = 0.0;
        if (l > 0) {
            for (int k = 0; k <= l; ++k) scale += std::fabs(A(i, k));
            if (scale == 0.0) {
                e[i] = A(i, l);
            } else {
                for (int k = 0; k <= l; ++k) { A(i, k) /= scale; h += A(i, k) * A(i, k); }
                double f = A(i, l);
                double g = f >= 0.0 ? -std::sqrt(h) : std::sqrt(h);
                e[i] = scale * g;
                h -= f * g;
                A(i, l) = f - g;
                f = 0.0;
                for (int j = 0; j <= l; ++j) {
                    g = 0.0;
                    for (int k = 0; k <= j; ++k) g += A(j, k) * A(i, k);
                    for (int k = j + 1; k <= l; ++k) g += A(k, j) * A(i, k);
                    e[j] = g / h;
                    f += e[j] * A(i, j);
                }
                const double hh = f / (h + h);
                for (int j = 0; j <= l; ++j) {
                    f = A(i, j);
                    e[j] = g = e[j] - hh * f;
                    for (int k = 0; k <= j; ++k) A(j, k) -= f * e[k] + g * A(i, k);
                }
            }
        } else {
            e[i] = A(i, l);
        }
        d[i] = h;
    }
    e[0] = 0.0;
    for (int i = 0; i < n; ++i) d[i] = A(i, i);
}

// number of eigenvalues of the symmetric tridiagonal (d, e) below x
static int sturm_count(const std::vector<double>& d, const std::vector<double>& e, int n, double x, double pivmin) {
    int cnt = 0;
    double q = d[0] - x;
    if (std::fabs(q) < pivmin) q = -pivmin;
    if (q < 0.0) ++cnt;
    for (int i = 1; i < n; ++i) {
        q = d[i] - x - e[i] * e[i] / q;
        if (std::fabs(q) < pivmin) q = -pivmin;
        if (q < 0.0) ++cnt;
    }
    return cnt;
}

static void tridiag_extremes(const std::vector<double>& d, const std::vector<double>& e, int n, double* lmin, double* lmax) {
    double lo = d[0], hi = d[0], emax = 0.0;
    for (int i = 0; i < n; ++i) {
        const double r = (i > 0 ? std::fabs(e[i]) : 0.0) + (i + 1 < n ? std::fabs(e[i + 1]) : 0.0);
        lo = std::min(lo, d[i] - r);
        hi = std::max(hi, d[i] + r);
        emax = std::max(emax, std::fabs(e[i]));
    }
    const double nrm = std::max(std::fabs(lo), std::fabs(hi));
    const double pivmin = std::max(2.2250738585072014e-308 * std::max(1.0, emax * emax), 1e-300);
    lo -= 2.0 * 2.220446049250313e-16 * nrm * n + 2.0 * pivmin;
    hi += 2.0 * 2.220446049250313e-16 * nrm * n + 2.0 * pivmin;
    auto kth = [&](int want) {        // eigenvalue with `want` eigenvalues below it (0-based, ascending)
        double a = lo, b = hi;
        for (int it = 0; it < 200; ++it) {
            const double m = 0.5 * (a + b);
            if (m <= a || m >= b) break;
            if (sturm_count(d, e, n, m, pivmin) > want) b = m; else a = m;
        }
        return 0.5 * (a + b);
    };
    *lmin = kth(0);
    *lmax = kth(n - 1);
}

// ---- general: reduction to upper Hessenberg form (a is n x n, row access a[i*n+j], destroyed) ---------------------
static void hessenberg_reduce(std::vector<double>& a, int n) {
    auto A = [&](int i, int j) -> double& { return a[(size_t)i * n + j]; };
    std::vector<double> ort(n, 0.0);
    for (int m = 1; m < n - 1; ++m) {
        double scale = 0.0;
        for (int i = m; i < n; ++i) scale += std::fabs(A(i, m - 1));
        if (scale == 0.0) continue;
        double h = 0.0;
        for (int i = n - 1; i >= m; --i) { ort[i] = A(i, m - 1) / scale; h += ort[i] * ort[i]; }
        const double g = ort[m] > 0.0 ? -std::sqrt(h) : std::sqrt(h);
        h -= ort[m] * g;
        ort[m] -= g;
        for (int j = m; j < n; ++j) {                       // (I - u u'/h) A
            double f = 0.0;
            for (int i = n - 1; i >= m; --i) f += ort[i] * A(i, j);
            f /= h;
            for (int i = m; i < n; ++i) A(i, j) -= f * ort[i];
        }
        for (int i = 0; i < n; ++i) {                       // A (I - u u'/h)
            double f = 0.0;
            for (int j = n - 1; j >= m; --j) f += ort[j] * A(i, j);
            f /= h;
            for (int j = m; j < n; ++j) A(i, j) -= f * ort[j];
        }
        ort[m] *= scale;
        A(m, m - 1) = scale * g;
        for (int i = m + 1; i < n; ++i) A(i, m - 1) = 0.0;
    }
}

// eigenvalues of an upper Hessenberg matrix by the double-shift QR iteration (eigenvalues only).  0 on success.
static int hessenberg_eigenvalues(std::vector<double>& a, int n, std::vector<double>& wr, std::vector<double>& wi) {
    auto A = [&](int i, int j) -> double& { return a[(size_t)i * n + j]; };
    wr.assign(n, 0.0);
    wi.assign(n, 0.0);
    double anorm = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = std::max(i - 1, 0); j < n; ++j) anorm += std::fabs(A(i, j));
    int nn = n - 1;
    double t = 0.0, p = 0.0, q = 0.0, r = 0.0, s = 0.0, w = 0.0, x = 0.0, y = 0.0, z = 0.0;
    const double EPS = 2.220446049250313e-16;
    while (nn >= 0) {
        int its = 0, l;
        do {
            for (l = nn; l >= 1; --l) {
                s = std::fabs(A(l - 1, l - 1)) + std::fabs(A(l, l));
                if (s == 0.0) s = anorm;
                if (std::fabs(A(l, l - 1)) <= EPS * s) { A(l, l - 1) = 0.0; break; }
            }
            x = A(nn, nn);
            if (l == nn) {                                   // one root
                wr[nn] = x + t; wi[nn--] = 0.0;
            } else {
                y = A(nn - 1, nn - 1);
                w = A(nn, nn - 1) * A(nn - 1, nn);
                if (l == nn - 1) {                           // two roots
                    p = 0.5 * (y - x);
                    q = p * p + w;
                    z = std::sqrt(std::fabs(q));
                    x += t;
                    if (q >= 0.0) {
                        z = p + (p >= 0.0 ? std::fabs(z) : -std::fabs(z));
                        wr[nn - 1] = wr[nn] = x + z;
                        if (z != 0.0) wr[nn] = x - w / z;
                        wi[nn - 1] = wi[nn] = 0.0;
                    } else {
                        wr[nn - 1] = wr[nn] = x + p;
                        wi[nn - 1] = -(wi[nn] = z);
                    }
                    nn -= 2;
                } else {
                    if (its == 60) return -1;
                    if (its == 10 || its == 20) {            // exceptional shift
                        t += x;
                        for (int i = 0; i <= nn; ++i) A(i, i) -= x;
                        s = std::fabs(A(nn, nn - 1)) + std::fabs(A(nn - 1, nn - 2));
                        y = x = 0.75 * s;
                        w = -0.4375 * s * s;
                    }
                    ++its;
                    int m;
                    for (m = nn - 2; m >= l; --m) {
                        z = A(m, m);
                        r = x - z;
                        s = y - z;
                        p = (r * s - w) / A(m + 1, m) + A(m, m + 1);
                        q = A(m + 1, m + 1) - z - r - s;
                        r = A(m + 2, m + 1);
                        s = std::fabs(p) + std::fabs(q) + std::fabs(r);
                        p /= s; q /= s; r /= s;
                        if (m == l) break;
                        const double u = std::fabs(A(m, m - 1)) * (std::fabs(q) + std::fabs(r));
                        const double v = std::fabs(p) * (std::fabs(A(m - 1, m - 1)) + std::fabs(z) + std::fabs(A(m + 1, m + 1)));
                        if (u <= EPS * v) break;
                    }
                    for (int i = m + 2; i <= nn; ++i) {
                        A(i, i - 2) = 0.0;
                        if (i != m + 2) A(i, i - 3) = 0.0;
                    }
                    for (int k = m; k <= nn - 1; ++k) {
                        if (k != m) {
                            p = A(k, k - 1);
                            q = A(k + 1, k - 1);
                            r = 0.0;
                            if (k != nn - 1) r = A(k + 2, k - 1);
                            if ((x = std::fabs(p) + std::fabs(q) + std::fabs(r)) != 0.0) { p /= x; q /= x; r /= x; }
                        }
                        const double sg = std::sqrt(p * p + q * q + r * r);
                        if ((s = (p >= 0.0 ? sg : -sg)) != 0.0) {
                            if (k == m) {
                                if (l != m) A(k, k - 1) = -A(k, k - 1);
                            } else {
                                A(k, k - 1) = -s * x;
                            }
                            p += s;
                            x = p / s; y = q / s; z = r / s;
                            q /= p; r /= p;
                            for (int j = k; j <= nn; ++j) {
                                p = A(k, j) + q * A(k + 1, j);
                                if (k != nn - 1) { p += r * A(k + 2, j); A(k + 2, j) -= p * z; }
                                A(k + 1, j) -= p * y;
                                A(k, j) -= p * x;
                            }
                            const int mmin = nn < k + 3 ? nn : k + 3;
                            for (int i = l; i <= mmin; ++i) {
                                p = x * A(i, k) + y * A(i, k + 1);
                                if (k != nn - 1) { p += z * A(i, k + 2); A(i, k + 2) -= p * r; }
                                A(i, k + 1) -= p * q;
                                A(i, k) -= p;
                            }
                        }
                    }
                }
            }
        } while (l < nn - 1);
    }
    return 0;
}

// ---- all minors k = 2..nmax ------------------------------------------------------------------------------------
struct ExtremeKey {
    uint64_t hash; int ld, nmax, kind;
    bool operator<(const ExtremeKey& o) const {
        if (hash != o.hash) return hash < o.hash;
        if (ld != o.ld) return ld < o.ld;
        if (nmax != o.nmax) return nmax < o.nmax;
        return kind < o.kind;
    }
};
static std::map<ExtremeKey, std::vector<double>> g_extreme_cache;   // [k] -> (min, max) pairs, k = 0..nmax
static std::mutex g_extreme_mutex;

// lead: leading ld x ld block of A_1, column-major (both triangles).  kind: 0 symmetric (min and max), 1 general (min).
// out: 2 * (nmax + 1) doubles, entries (2k, 2k+1) = (min, max) eigenvalue of the k x k minor; max = NaN for kind 1.
int minor_extremes(const std::vector<double>& lead, int ld, int nmax, int kind, std::vector<double>& out) {
    uint64_t hsh = 1469598103934665603ULL;
    const unsigned char* bytes = reinterpret_cast<const unsigned char*>(lead.data());
    for (size_t i = 0; i < lead.size() * 8; ++i) { hsh ^= bytes[i]; hsh *= 1099511628211ULL; }
    const ExtremeKey key{hsh, ld, nmax, kind};
    {
        std::lock_guard<std::mutex> lock(g_extreme_mutex);
        auto it = g_extreme_cache.find(key);
        if (it != g_extreme_cache.end()) { out = it->second; return 0; }
    }
    out.assign(2 * (size_t)(nmax + 1), NAN);
    // structure of the block: bandwidths decide whether a reduction is needed at all
    int lower_bw = 0, upper_bw = 0;
    for (int j = 0; j < nmax; ++j)
        for (int i = 0; i < nmax; ++i)
            if (lead[(size_t)j * ld + i] != 0.0) { lower_bw = std::max(lower_bw, i - j); upper_bw = std::max(upper_bw, j - i); }
    std::atomic<int> next(nmax), failed(0), complex_k(0);
    auto work = [&]() {
        std::vector<double> a, d, e, wr, wi;
        for (;;) {
            const int k = next.fetch_sub(1);              // largest minors first: they cost the most
            if (k < 1) break;
            if (kind == 0) {
                if (lower_bw <= 1) {
                    d.assign(k, 0.0); e.assign(k, 0.0);
                    for (int i = 0; i < k; ++i) { d[i] = lead[(size_t)i * ld + i]; if (i > 0) e[i] = lead[(size_t)(i - 1) * ld + i]; }
                } else {
                    a.assign((size_t)k * k, 0.0);
                    for (int j = 0; j < k; ++j)
                        for (int i = 0; i < k; ++i) a[(size_t)j * k + i] = lead[(size_t)j * ld + i];
                    tridiagonalize(a, k, d, e);
                }
                tridiag_extremes(d, e, k, &out[2 * k], &out[2 * k + 1]);
            } else {
                a.assign((size_t)k * k, 0.0);
                for (int i = 0; i < k; ++i)
                    for (int j = 0; j < k; ++j) a[(size_t)i * k + j] = lead[(size_t)j * ld + i];   // row access
                if (lower_bw > 1) hessenberg_reduce(a, k);
                if (hessenberg_eigenvalues(a, k, wr, wi) != 0) { failed = 1; continue; }
                double mn = wr[0], imax = 0.0;
                for (int i = 0; i < k; ++i) { mn = std::min(mn, wr[i]); imax = std::max(imax, std::fabs(wi[i])); }
                if (imax > 0.0) { int zero = 0; complex_k.compare_exchange_strong(zero, k); }
                out[2 * k] = mn;
            }
        }
    };
    unsigned nthr = std::thread::hardware_concurrency();
    nthr = std::max(1u, std::min(nthr ? nthr : 4u, (unsigned)std::max(1, nmax / 4)));
    std::vector<std::thread> pool;
    for (unsigned i = 1; i < nthr; ++i) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    if (failed) return set_error(TK_ESTATE, "QR iteration on a minor of A_1 did not converge");
    if (complex_k)       // Julia: minimum(eigvals(...)) of a Complex vector has no method (eigenvalues.jl:344-350)
        return set_error(TK_EUNSUPPORTED, "the %d x %d leading minor of A_1 has complex eigenvalues: the reference's spectral data (eigenvalues.jl:344-350) is undefined", (int)complex_k, (int)complex_k);
    std::lock_guard<std::mutex> lock(g_extreme_mutex);
    g_extreme_cache[key] = out;
    return 0;
}

}  // namespace tk

extern "C" {

int tk_minor_extremes(const double* lead, int32_t ld, int32_t nmax, int32_t general, double* out) {
    if (!lead || !out || ld < 1 || nmax < 1 || nmax > ld) return tk::set_error(TK_EINVAL, "bad arguments");
    std::vector<double> a(lead, lead + (size_t)ld * ld), res;
    const int rc = tk::minor_extremes(a, ld, nmax, general ? 1 : 0, res);
    if (rc) return rc;
    std::memcpy(out, res.data(), 8 * res.size());
    return 0;
}

}  // extern "C"
