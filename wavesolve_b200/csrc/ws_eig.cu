// K5-K7, K9, K10: Krylov eigensolver for the shift-inverted pencil, OP = (A - sigma B)^-1 B.
//
//  kind 0 (scalar, B SPD):  Arnoldi/Lanczos in the B inner product with full (CGS2)
//          re-orthogonalisation and thick restart -- what ARPACK's dsaupd/dseupd mode 3 does
//          behind eigsh(A, M=B, k, which='SA', sigma) at fe_solver.py:328.  The dual basis
//          Z = B V is carried along, so a step costs one solve + one SpMV.
//  kind 1 (vector, B indefinite):  Euclidean Arnoldi with thick (Krylov-Schur style) restart on
//          the non-symmetric OP -- the matrix-free equivalent of eigs(spsolve(A - sigma B, B),
//          k, which='SR') at fe_solver.py:398-400.  OP x = T M'^-1 T^T (B x) with the
//          quasi-definite M' (ws_assemble_vector_qd).
//
// Device side: the basis lives in HBM as (ncv+1) rows of length N; one Arnoldi step is a
// fixed launch sequence with every coefficient kept in device memory (no host round trip
// inside a sweep).  Host side: the (ncv+1) x ncv projected matrix is copied back once per
// sweep and its eigen-decomposition (Jacobi for the symmetric case, Hessenberg + shifted QR
// otherwise), the convergence test and the restart rotation are computed in plain C++.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <memory>
#include <numeric>
#include <thread>

#include "ws_solver.cuh"

namespace ws {

void spmv_launch(const ws_pattern* p, const double* vals, const double* x, double* y, cudaStream_t st);
void spmv_launch2(const ws_pattern* p, const double* vals, const double* x, double* y, cudaStream_t st);

// =============================================================================================
// host dense eigen-solvers (n <= MAX_NCV)
// =============================================================================================
// symmetric, small (n <= 48): cyclic Jacobi.  A row-major n x n (destroyed), w eigenvalues,
// V[i*n+j] = component i of vector j
static void jacobi_small(int n, std::vector<double>& A, std::vector<double>& w, std::vector<double>& V) {
    V.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
    auto a = [&](int i, int j) -> double& { return A[(size_t)i * n + j]; };
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; ++i) {
            diag += a(i, i) * a(i, i);
            for (int j = i + 1; j < n; ++j) off += a(i, j) * a(i, j);
        }
        if (off <= 1e-34 * (diag + off) || off == 0.0) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = a(p, q);
                if (apq == 0.0) continue;
                const double theta = (a(q, q) - a(p, p)) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = a(k, p), akq = a(k, q);
                    a(k, p) = c * akp - s * akq;
                    a(k, q) = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = a(p, k), aqk = a(q, k);
                    a(p, k) = c * apk - s * aqk;
                    a(q, k) = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
                    V[(size_t)k * n + p] = c * vkp - s * vkq;
                    V[(size_t)k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    w.resize(n);
    for (int i = 0; i < n; ++i) w[i] = a(i, i);
}

// symmetric, wide bases (ncv up to 256 for Nmax ~ 100: Jacobi would cost ~10^8 flops per restart):
// Householder reduction to tridiagonal form with the transformation accumulated in place, then the
// implicit-shift QL iteration on the tridiagonal (the classical tred2 / tql2 scheme), O(n^3) once.
static bool tridiag_ql(int n, std::vector<double>& A, std::vector<double>& w, std::vector<double>& V) {
    V = A;
    auto z = [&](int i, int j) -> double& { return V[(size_t)i * n + j]; };
    std::vector<double> d(n, 0.0), e(n, 0.0);
    for (int i = n - 1; i >= 1; --i) {
        const int l = i;                       // columns 0..l-1 of row i are eliminated down to one entry
        double h = 0.0, scale = 0.0;
        if (l > 1) {
            for (int k = 0; k < l; ++k) scale += std::fabs(z(i, k));
            if (scale == 0.0) {
                e[i] = z(i, l - 1);
            } else {
                for (int k = 0; k < l; ++k) {
                    z(i, k) /= scale;
                    h += z(i, k) * z(i, k);
                }
                double f = z(i, l - 1);
                double g = f >= 0.0 ? -std::sqrt(h) : std::sqrt(h);
                e[i] = scale * g;
                h -= f * g;
                z(i, l - 1) = f - g;
                f = 0.0;
                for (int j = 0; j < l; ++j) {
                    z(j, i) = z(i, j) / h;
                    g = 0.0;
                    for (int k = 0; k <= j; ++k) g += z(j, k) * z(i, k);
                    for (int k = j + 1; k < l; ++k) g += z(k, j) * z(i, k);
                    e[j] = g / h;
                    f += e[j] * z(i, j);
                }
                const double hh = f / (h + h);
                for (int j = 0; j < l; ++j) {
                    f = z(i, j);
                    e[j] = g = e[j] - hh * f;
                    for (int k = 0; k <= j; ++k) z(j, k) -= f * e[k] + g * z(i, k);
                }
            }
        } else {
            e[i] = z(i, l - 1);
        }
        d[i] = h;
    }
    d[0] = 0.0;
    e[0] = 0.0;
    for (int i = 0; i < n; ++i) {
        if (d[i] != 0.0)
            for (int j = 0; j < i; ++j) {
                double g = 0.0;
                for (int k = 0; k < i; ++k) g += z(i, k) * z(k, j);
                for (int k = 0; k < i; ++k) z(k, j) -= g * z(k, i);
            }
        d[i] = z(i, i);
        z(i, i) = 1.0;
        for (int j = 0; j < i; ++j) z(j, i) = z(i, j) = 0.0;
    }
    // implicit QL on (d, e), rotations applied to the columns of V
    for (int i = 1; i < n; ++i) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    for (int l = 0; l < n; ++l) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; ++m) {
                const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
                if (std::fabs(e[m]) <= 2.220446049250313e-16 * dd) break;
            }
            if (m != l) {
                if (++iter > 120) return false;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = std::hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; --i) {
                    double f = s * e[i];
                    const double b = c * e[i];
                    e[i + 1] = (r = std::hypot(f, g));
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[m] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    d[i + 1] = g + (p = s * r);
                    g = c * r - b;
                    for (int k = 0; k < n; ++k) {
                        f = z(k, i + 1);
                        z(k, i + 1) = s * z(k, i) + c * f;
                        z(k, i) = c * z(k, i) - s * f;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
    w = d;
    return true;
}

void jacobi_sym(int n, std::vector<double>& A, std::vector<double>& w, std::vector<double>& V) {
    if (n > 48) {
        std::vector<double> A2 = A;
        if (tridiag_ql(n, A2, w, V)) return;
    }
    jacobi_small(n, A, w, V);
}

// general real matrix: Householder reduction to Hessenberg form + implicit double-shift QR with
// accumulated transformations and back-substitution (the classical EISPACK orthes/hqr2 scheme).
// H row-major n x n (destroyed).  wr/wi eigenvalues; V[i*n+j]: real eigenvalue j -> column j is
// its eigenvector; complex pair (j, j+1), wi[j] > 0 -> columns j, j+1 are the real and imaginary
// parts of the eigenvector of wr[j] + i wi[j].  Returns false if the QR iteration stalls.
static void cdiv(double xr, double xi, double yr, double yi, double& cr, double& ci) {
    double r, d;
    if (std::fabs(yr) > std::fabs(yi)) {
        r = yi / yr; d = yr + r * yi;
        cr = (xr + r * xi) / d; ci = (xi - r * xr) / d;
    } else {
        r = yr / yi; d = yi + r * yr;
        cr = (r * xr + xi) / d; ci = (r * xi - xr) / d;
    }
}

bool eig_general(int nn, std::vector<double>& Hm, std::vector<double>& wr, std::vector<double>& wi,
                 std::vector<double>& Vm) {
    auto H = [&](int i, int j) -> double& { return Hm[(size_t)i * nn + j]; };
    Vm.assign((size_t)nn * nn, 0.0);
    auto V = [&](int i, int j) -> double& { return Vm[(size_t)i * nn + j]; };
    wr.assign(nn, 0.0);
    wi.assign(nn, 0.0);
    std::vector<double> ort(nn, 0.0);
    const int low = 0, high = nn - 1;
    // ---- reduction to Hessenberg form
    for (int m = low + 1; m <= high - 1; ++m) {
        double scale = 0.0;
        for (int i = m; i <= high; ++i) scale += std::fabs(H(i, m - 1));
        if (scale != 0.0) {
            double h = 0.0;
            for (int i = high; i >= m; --i) {
                ort[i] = H(i, m - 1) / scale;
                h += ort[i] * ort[i];
            }
            double g = std::sqrt(h);
            if (ort[m] > 0) g = -g;
            h -= ort[m] * g;
            ort[m] -= g;
            for (int j = m; j < nn; ++j) {
                double f = 0.0;
                for (int i = high; i >= m; --i) f += ort[i] * H(i, j);
                f /= h;
                for (int i = m; i <= high; ++i) H(i, j) -= f * ort[i];
            }
            for (int i = 0; i <= high; ++i) {
                double f = 0.0;
                for (int j = high; j >= m; --j) f += ort[j] * H(i, j);
                f /= h;
                for (int j = m; j <= high; ++j) H(i, j) -= f * ort[j];
            }
            ort[m] = scale * ort[m];
            H(m, m - 1) = scale * g;
        }
    }
    for (int i = 0; i < nn; ++i) V(i, i) = 1.0;
    for (int m = high - 1; m >= low + 1; --m) {
        if (H(m, m - 1) != 0.0) {
            for (int i = m + 1; i <= high; ++i) ort[i] = H(i, m - 1);
            for (int j = m; j <= high; ++j) {
                double g = 0.0;
                for (int i = m; i <= high; ++i) g += ort[i] * V(i, j);
                g = (g / ort[m]) / H(m, m - 1);
                for (int i = m; i <= high; ++i) V(i, j) += g * ort[i];
            }
        }
    }
    // ---- QR iteration
    int n = nn - 1;
    const double eps = std::pow(2.0, -52.0);
    double exshift = 0.0, p = 0, q = 0, r = 0, s = 0, z = 0, t, w, x, y;
    double norm = 0.0;
    for (int i = 0; i < nn; ++i)
        for (int j = std::max(i - 1, 0); j < nn; ++j) norm += std::fabs(H(i, j));
    int iter = 0, total = 0;
    while (n >= low) {
        int l = n;
        while (l > low) {
            s = std::fabs(H(l - 1, l - 1)) + std::fabs(H(l, l));
            if (s == 0.0) s = norm;
            if (std::fabs(H(l, l - 1)) < eps * s) break;
            l--;
        }
        if (l == n) {
            H(n, n) += exshift;
            wr[n] = H(n, n);
            wi[n] = 0.0;
            n--;
            iter = 0;
        } else if (l == n - 1) {
            w = H(n, n - 1) * H(n - 1, n);
            p = (H(n - 1, n - 1) - H(n, n)) / 2.0;
            q = p * p + w;
            z = std::sqrt(std::fabs(q));
            H(n, n) += exshift;
            H(n - 1, n - 1) += exshift;
            x = H(n, n);
            if (q >= 0) {
                z = (p >= 0) ? p + z : p - z;
                wr[n - 1] = x + z;
                wr[n] = wr[n - 1];
                if (z != 0.0) wr[n] = x - w / z;
                wi[n - 1] = 0.0;
                wi[n] = 0.0;
                x = H(n, n - 1);
                s = std::fabs(x) + std::fabs(z);
                p = x / s;
                q = z / s;
                r = std::sqrt(p * p + q * q);
                p /= r;
                q /= r;
                for (int j = n - 1; j < nn; ++j) {
                    z = H(n - 1, j);
                    H(n - 1, j) = q * z + p * H(n, j);
                    H(n, j) = q * H(n, j) - p * z;
                }
                for (int i = 0; i <= n; ++i) {
                    z = H(i, n - 1);
                    H(i, n - 1) = q * z + p * H(i, n);
                    H(i, n) = q * H(i, n) - p * z;
                }
                for (int i = low; i <= high; ++i) {
                    z = V(i, n - 1);
                    V(i, n - 1) = q * z + p * V(i, n);
                    V(i, n) = q * V(i, n) - p * z;
                }
            } else {
                wr[n - 1] = x + p;
                wr[n] = x + p;
                wi[n - 1] = z;
                wi[n] = -z;
            }
            n -= 2;
            iter = 0;
        } else {
            x = H(n, n);
            y = 0.0;
            w = 0.0;
            if (l < n) {
                y = H(n - 1, n - 1);
                w = H(n, n - 1) * H(n - 1, n);
            }
            if (iter == 10) {
                exshift += x;
                for (int i = low; i <= n; ++i) H(i, i) -= x;
                s = std::fabs(H(n, n - 1)) + std::fabs(H(n - 1, n - 2));
                x = y = 0.75 * s;
                w = -0.4375 * s * s;
            }
            if (iter == 30) {
                s = (y - x) / 2.0;
                s = s * s + w;
                if (s > 0) {
                    s = std::sqrt(s);
                    if (y < x) s = -s;
                    s = x - w / ((y - x) / 2.0 + s);
                    for (int i = low; i <= n; ++i) H(i, i) -= s;
                    exshift += s;
                    x = y = w = 0.964;
                }
            }
            iter++;
            if (++total > 3000 + 60 * nn) return false;
            int m = n - 2;
            while (m >= l) {
                z = H(m, m);
                r = x - z;
                s = y - z;
                p = (r * s - w) / H(m + 1, m) + H(m, m + 1);
                q = H(m + 1, m + 1) - z - r - s;
                r = H(m + 2, m + 1);
                s = std::fabs(p) + std::fabs(q) + std::fabs(r);
                p /= s;
                q /= s;
                r /= s;
                if (m == l) break;
                if (std::fabs(H(m, m - 1)) * (std::fabs(q) + std::fabs(r)) <
                    eps * (std::fabs(p) * (std::fabs(H(m - 1, m - 1)) + std::fabs(z) + std::fabs(H(m + 1, m + 1)))))
                    break;
                m--;
            }
            for (int i = m + 2; i <= n; ++i) {
                H(i, i - 2) = 0.0;
                if (i > m + 2) H(i, i - 3) = 0.0;
            }
            for (int k = m; k <= n - 1; ++k) {
                const bool notlast = (k != n - 1);
                if (k != m) {
                    p = H(k, k - 1);
                    q = H(k + 1, k - 1);
                    r = notlast ? H(k + 2, k - 1) : 0.0;
                    x = std::fabs(p) + std::fabs(q) + std::fabs(r);
                    if (x == 0.0) continue;
                    p /= x;
                    q /= x;
                    r /= x;
                }
                s = std::sqrt(p * p + q * q + r * r);
                if (p < 0) s = -s;
                if (s != 0) {
                    if (k != m) H(k, k - 1) = -s * x;
                    else if (l != m) H(k, k - 1) = -H(k, k - 1);
                    p += s;
                    x = p / s;
                    y = q / s;
                    z = r / s;
                    q /= p;
                    r /= p;
                    for (int j = k; j < nn; ++j) {
                        p = H(k, j) + q * H(k + 1, j);
                        if (notlast) {
                            p += r * H(k + 2, j);
                            H(k + 2, j) -= p * z;
                        }
                        H(k, j) -= p * x;
                        H(k + 1, j) -= p * y;
                    }
                    for (int i = 0; i <= std::min(n, k + 3); ++i) {
                        p = x * H(i, k) + y * H(i, k + 1);
                        if (notlast) {
                            p += z * H(i, k + 2);
                            H(i, k + 2) -= p * r;
                        }
                        H(i, k) -= p;
                        H(i, k + 1) -= p * q;
                    }
                    for (int i = low; i <= high; ++i) {
                        p = x * V(i, k) + y * V(i, k + 1);
                        if (notlast) {
                            p += z * V(i, k + 2);
                            V(i, k + 2) -= p * r;
                        }
                        V(i, k) -= p;
                        V(i, k + 1) -= p * q;
                    }
                }
            }
        }
    }
    if (norm == 0.0) return true;
    // ---- back-substitution: eigenvectors of the quasi-triangular form
    for (n = nn - 1; n >= 0; --n) {
        p = wr[n];
        q = wi[n];
        if (q == 0) {
            int l = n;
            H(n, n) = 1.0;
            for (int i = n - 1; i >= 0; --i) {
                w = H(i, i) - p;
                r = 0.0;
                for (int j = l; j <= n; ++j) r += H(i, j) * H(j, n);
                if (wi[i] < 0.0) {
                    z = w;
                    s = r;
                } else {
                    l = i;
                    if (wi[i] == 0.0) {
                        if (w != 0.0) H(i, n) = -r / w;
                        else H(i, n) = -r / (eps * norm);
                    } else {
                        x = H(i, i + 1);
                        y = H(i + 1, i);
                        q = (wr[i] - p) * (wr[i] - p) + wi[i] * wi[i];
                        t = (x * s - z * r) / q;
                        H(i, n) = t;
                        if (std::fabs(x) > std::fabs(z)) H(i + 1, n) = (-r - w * t) / x;
                        else H(i + 1, n) = (-s - y * t) / z;
                    }
                    t = std::fabs(H(i, n));
                    if ((eps * t) * t > 1)
                        for (int j = i; j <= n; ++j) H(j, n) /= t;
                }
            }
        } else if (q < 0) {
            int l = n - 1;
            double cr, ci;
            if (std::fabs(H(n, n - 1)) > std::fabs(H(n - 1, n))) {
                H(n - 1, n - 1) = q / H(n, n - 1);
                H(n - 1, n) = -(H(n, n) - p) / H(n, n - 1);
            } else {
                cdiv(0.0, -H(n - 1, n), H(n - 1, n - 1) - p, q, cr, ci);
                H(n - 1, n - 1) = cr;
                H(n - 1, n) = ci;
            }
            H(n, n - 1) = 0.0;
            H(n, n) = 1.0;
            for (int i = n - 2; i >= 0; --i) {
                double ra = 0.0, sa = 0.0, vr, vi;
                for (int j = l; j <= n; ++j) {
                    ra += H(i, j) * H(j, n - 1);
                    sa += H(i, j) * H(j, n);
                }
                w = H(i, i) - p;
                if (wi[i] < 0.0) {
                    z = w;
                    r = ra;
                    s = sa;
                } else {
                    l = i;
                    if (wi[i] == 0) {
                        cdiv(-ra, -sa, w, q, cr, ci);
                        H(i, n - 1) = cr;
                        H(i, n) = ci;
                    } else {
                        x = H(i, i + 1);
                        y = H(i + 1, i);
                        vr = (wr[i] - p) * (wr[i] - p) + wi[i] * wi[i] - q * q;
                        vi = (wr[i] - p) * 2.0 * q;
                        if (vr == 0.0 && vi == 0.0)
                            vr = eps * norm * (std::fabs(w) + std::fabs(q) + std::fabs(x) + std::fabs(y) + std::fabs(z));
                        cdiv(x * r - z * ra + q * sa, x * s - z * sa - q * ra, vr, vi, cr, ci);
                        H(i, n - 1) = cr;
                        H(i, n) = ci;
                        if (std::fabs(x) > (std::fabs(z) + std::fabs(q))) {
                            H(i + 1, n - 1) = (-ra - w * H(i, n - 1) + q * H(i, n)) / x;
                            H(i + 1, n) = (-sa - w * H(i, n) - q * H(i, n - 1)) / x;
                        } else {
                            cdiv(-r - y * H(i, n - 1), -s - y * H(i, n), z, q, cr, ci);
                            H(i + 1, n - 1) = cr;
                            H(i + 1, n) = ci;
                        }
                    }
                    t = std::max(std::fabs(H(i, n - 1)), std::fabs(H(i, n)));
                    if ((eps * t) * t > 1)
                        for (int j = i; j <= n; ++j) {
                            H(j, n - 1) /= t;
                            H(j, n) /= t;
                        }
                }
            }
        }
    }
    for (int j = nn - 1; j >= low; --j)
        for (int i = low; i <= high; ++i) {
            z = 0.0;
            for (int k = low; k <= std::min(j, high); ++k) z += V(i, k) * H(k, j);
            V(i, j) = z;
        }
    return true;
}

// =============================================================================================
// device kernels on the Krylov basis (rows of length N)
// =============================================================================================
constexpr int DOT_CHUNK = 2048;
constexpr int MAXV = 64;        // basis rows the FUSED CGS kernel and the block-2 kernels handle (static buffers)
constexpr int MAX_NCV = 256;    // largest Krylov basis: Nmax = 100 (lantern-demo.ipynb cells 2, 4) needs ncv = 204

// partial[c*nvec + i] = sum over chunk c of Z_i[n] * u[n];  8 warps, warp w takes vectors w, w+8, ...
// Optional fused application of T (vector problem: w = T x, x the output of the solve sweeps): with tf.tin != nullptr the
// chunk of u is formed on the fly as u[n] = tin[n] - (tin[Ntt + hi] - tin[Ntt + lo]) / len[n] for the edge rows (what
// k_vector_T mode 0 computes) and also written to u -- one launch and one pass over the vector less per Arnoldi step.
struct FusedT {
    const double* tin = nullptr;
    const int2* edges = nullptr;
    const double* elen = nullptr;
    int64_t Ntt = 0;
    double* wout = nullptr;
};
__global__ void __launch_bounds__(256) k_dots_partial(const double* __restrict__ Z, int64_t ldz, int nvec,
                                                      const double* __restrict__ u, int64_t N,
                                                      double* __restrict__ partial, int chunk, FusedT tf) {
    __shared__ double us[DOT_CHUNK];
    const int64_t n0 = (int64_t)blockIdx.x * chunk;
    const int len = (int)min((int64_t)chunk, N - n0);
    if (tf.tin) {
        for (int i = threadIdx.x; i < chunk; i += 256) {
            double v = 0.0;
            if (i < len) {
                const int64_t n = n0 + i;
                v = tf.tin[n];
                if (n < tf.Ntt) {
                    const int2 e = tf.edges[n];
                    v -= (tf.tin[tf.Ntt + e.y] - tf.tin[tf.Ntt + e.x]) / tf.elen[n];
                }
                if (blockIdx.y == 0) tf.wout[n] = v;
            }
            us[i] = v;
        }
    } else {
        for (int i = threadIdx.x; i < chunk; i += 256) us[i] = (i < len) ? u[n0 + i] : 0.0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // gridDim.y groups of basis vectors: a wide basis on a small problem (Nmax = 80 at N = 100k: 164 vectors, 50 chunks)
    // would otherwise leave two thirds of the SMs idle while 50 CTAs walk 164 vectors each
    const int per = (nvec + gridDim.y - 1) / gridDim.y;
    const int v0 = blockIdx.y * per, v1 = min(nvec, v0 + per);
    for (int v = v0 + warp; v < v1; v += 8) {
        const double* z = Z + (size_t)v * ldz + n0;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int i = lane;
        for (; i + 96 < len; i += 128) {
            a0 += z[i] * us[i];
            a1 += z[i + 32] * us[i + 32];
            a2 += z[i + 64] * us[i + 64];
            a3 += z[i + 96] * us[i + 96];
        }
        for (; i < len; i += 32) a0 += z[i] * us[i];
        double acc = (a0 + a1) + (a2 + a3);
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) partial[(size_t)blockIdx.x * nvec + v] = acc;
    }
}

// h[i] = sum_c partial[c][i]; one warp per vector, fixed lane/shuffle order (deterministic);
// hacc[i] += h[i] when hacc != nullptr
__global__ void __launch_bounds__(32) k_dots_reduce(const double* __restrict__ partial, int nchunks, int nvec,
                                                    double* __restrict__ h, double* __restrict__ hacc) {
    const int i = blockIdx.x, lane = threadIdx.x;
    double s = 0.0;
    for (int c = lane; c < nchunks; c += 32) s += partial[(size_t)c * nvec + i];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        h[i] = s;
        if (hacc) hacc[i] += s;
    }
}

// w -= sum_i h[i] V_i
__global__ void __launch_bounds__(256) k_update(double* __restrict__ w, const double* __restrict__ V, int64_t ldv,
                                                int nvec, const double* __restrict__ h, int64_t N) {
    extern __shared__ double hs[];             // [nvec]
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) hs[i] = h[i];
    __syncthreads();
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= N) return;
    double acc = w[n];
    for (int i = 0; i < nvec; ++i) acc -= hs[i] * V[(size_t)i * ldv + n];
    w[n] = acc;
}

// CGS2, middle of the step: w -= V h (update of pass 1) and the partial inner products V^T w of pass 2
// in ONE pass over the basis.  A CTA owns DOT_CHUNK rows and walks them in tiles of 256: the tile of
// every basis vector is staged in shared memory once, used to update the 256 entries of w and then,
// with the updated entries, for the partial dots.  partial has the layout of k_dots_partial.
constexpr int UD_TILE = 256;
__global__ void __launch_bounds__(256) k_update_dots(double* __restrict__ w, const double* __restrict__ V, int64_t ldv, int nvec,
                                                     const double* __restrict__ h, int64_t N, double* __restrict__ partial, int chunk,
                                                     double* __restrict__ pnorm) {
    extern __shared__ double uds[];            // [nvec][UD_TILE] basis tile, then [UD_TILE] updated w
    __shared__ double hs[MAXV];
    __shared__ double sqw[8];
    double sq = 0.0;                           // sum of the squares of this thread's updated entries (pnorm != nullptr)
    double* Vs = uds;
    double* ws_ = uds + (size_t)nvec * UD_TILE;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < nvec) hs[tid] = h[tid];
    const int64_t c0 = (int64_t)blockIdx.x * chunk;
    double dsum[MAXV / 8];
#pragma unroll
    for (int i = 0; i < MAXV / 8; ++i) dsum[i] = 0.0;
    for (int t0 = 0; t0 < chunk; t0 += UD_TILE) {
        const int64_t n0 = c0 + t0;
        if (n0 >= N) break;
        const int len = (int)min((int64_t)UD_TILE, N - n0);
        __syncthreads();                       // previous tile fully consumed (and hs visible)
        for (int v = warp; v < nvec; v += 8) {
            const double* src = V + (size_t)v * ldv + n0;
#pragma unroll
            for (int k = 0; k < UD_TILE / 32; ++k) {
                const int i = lane + 32 * k;
                Vs[v * UD_TILE + i] = i < len ? src[i] : 0.0;
            }
        }
        __syncthreads();
        double acc = 0.0;
        if (tid < len) {
            acc = w[n0 + tid];
            for (int v = 0; v < nvec; ++v) acc -= hs[v] * Vs[v * UD_TILE + tid];
            w[n0 + tid] = acc;
        }
        sq += acc * acc;
        ws_[tid] = acc;
        __syncthreads();
        int slot = 0;
        for (int v = warp; v < nvec; v += 8, ++slot) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < UD_TILE / 32; ++k) a += Vs[v * UD_TILE + lane + 32 * k] * ws_[lane + 32 * k];
            dsum[slot] += a;
        }
    }
    int slot = 0;
    for (int v = warp; v < nvec; v += 8, ++slot) {
        double a = dsum[slot];
#pragma unroll
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) partial[(size_t)blockIdx.x * nvec + v] = a;
    }
    if (pnorm) {
        // ||w'||^2 of this chunk (fixed shuffle / warp order: deterministic), for beta^2 = ||w'||^2 - ||h2||^2
#pragma unroll
        for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (lane == 0) sqw[warp] = sq;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int k = 0; k < 8; ++k) t += sqw[k];
            pnorm[blockIdx.x] = t;
        }
    }
}

// second reduction of the fused CGS2 step: h2[v] = sum_c partial[c][v] (also accumulated into the H column), and
// nrm[0] = beta^2 = ||w'||^2 - ||h2||^2 -- the squared norm of w'' = w' - V h2 by Pythagoras (V orthonormal to rounding,
// ||h2|| ~ 1e-16 ||w'||: no cancellation), so that the last update and the normalisation are ONE pass (k_update_normalize)
__global__ void __launch_bounds__(256) k_dots_reduce_beta(const double* __restrict__ partial, const double* __restrict__ pnorm,
                                                          int nchunks, int nvec, double* __restrict__ h, double* __restrict__ hacc,
                                                          double* __restrict__ nrm) {
    __shared__ double h2sq[MAXV];
    __shared__ double pw[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int v = warp; v < nvec; v += 8) {
        double sacc = 0.0;
        for (int c = lane; c < nchunks; c += 32) sacc += partial[(size_t)c * nvec + v];
#pragma unroll
        for (int o = 16; o; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
        if (lane == 0) {
            h[v] = sacc;
            if (hacc) hacc[v] += sacc;
            h2sq[v] = sacc * sacc;
        }
    }
    double pn = 0.0;
    for (int c = threadIdx.x; c < nchunks; c += 256) pn += pnorm[c];
#pragma unroll
    for (int o = 16; o; o >>= 1) pn += __shfl_xor_sync(0xffffffffu, pn, o);
    if (lane == 0) pw[warp] = pn;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0, hh = 0.0;
        for (int k = 0; k < 8; ++k) t += pw[k];
        for (int v = 0; v < nvec; ++v) hh += h2sq[v];
        nrm[0] = t - hh;
    }
}

// v_next = (w - V h) / beta, beta = sqrt(nrm[0]) also written to the H slot: last CGS2 update + normalisation in one pass
__global__ void __launch_bounds__(256) k_update_normalize(const double* __restrict__ w, const double* __restrict__ V, int64_t ldv,
                                                          int nvec, const double* __restrict__ h, int64_t N,
                                                          const double* __restrict__ nrm, double* __restrict__ hslot,
                                                          double* __restrict__ vnext) {
    extern __shared__ double hs[];             // [nvec]
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) hs[i] = h[i];
    __syncthreads();
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const double b2 = nrm[0];
    const double beta = b2 > 0.0 ? sqrt(b2) : 0.0;
    if (n == 0 && hslot) *hslot = beta;
    if (n >= N) return;
    const double inv = beta > 0.0 ? 1.0 / beta : 0.0;
    double acc = w[n];
    for (int i = 0; i < nvec; ++i) acc -= hs[i] * V[(size_t)i * ldv + n];
    vnext[n] = acc * inv;
}

// two vectors against the same basis in one pass over it (block Arnoldi):
// partial[c*2*nvec + 2*i + r] = sum over chunk c of Z_i[n] * u_r[n]
__global__ void __launch_bounds__(256) k_dots_partial2(const double* __restrict__ Z, int64_t ldz, int nvec,
                                                       const double* __restrict__ u0, const double* __restrict__ u1, int64_t N,
                                                       double* __restrict__ partial) {
    __shared__ double us[2][DOT_CHUNK];
    const int64_t n0 = (int64_t)blockIdx.x * DOT_CHUNK;
    const int len = (int)min((int64_t)DOT_CHUNK, N - n0);
    for (int i = threadIdx.x; i < DOT_CHUNK; i += 256) {
        us[0][i] = (i < len) ? u0[n0 + i] : 0.0;
        us[1][i] = (i < len) ? u1[n0 + i] : 0.0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int v = warp; v < nvec; v += 8) {
        const double* z = Z + (size_t)v * ldz + n0;
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
        int i = lane;
        for (; i + 32 < len; i += 64) {
            const double z0 = z[i], z1 = z[i + 32];
            a0 += z0 * us[0][i];
            b0 += z0 * us[1][i];
            a1 += z1 * us[0][i + 32];
            b1 += z1 * us[1][i + 32];
        }
        for (; i < len; i += 32) {
            const double z0 = z[i];
            a0 += z0 * us[0][i];
            b0 += z0 * us[1][i];
        }
        double sa = a0 + a1, sb = b0 + b1;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
            sb += __shfl_xor_sync(0xffffffffu, sb, o);
        }
        if (lane == 0) {
            partial[((size_t)blockIdx.x * nvec + v) * 2] = sa;
            partial[((size_t)blockIdx.x * nvec + v) * 2 + 1] = sb;
        }
    }
}

// h[2*i + r] = sum_c partial[c][i][r];  hacc_r[i] += h[2*i + r]
__global__ void __launch_bounds__(32) k_dots_reduce2(const double* __restrict__ partial, int nchunks, int nvec,
                                                     double* __restrict__ h, double* __restrict__ hacc0, double* __restrict__ hacc1) {
    const int i = blockIdx.x >> 1, r = blockIdx.x & 1, lane = threadIdx.x;
    double s = 0.0;
    for (int c = lane; c < nchunks; c += 32) s += partial[((size_t)c * nvec + i) * 2 + r];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        h[2 * i + r] = s;
        double* hacc = r ? hacc1 : hacc0;
        if (hacc) hacc[i] += s;
    }
}

// w_r -= sum_i h[2*i + r] V_i, r = 0, 1: one pass over the basis for both
__global__ void __launch_bounds__(256) k_update2(double* __restrict__ w0, double* __restrict__ w1, const double* __restrict__ V,
                                                 int64_t ldv, int nvec, const double* __restrict__ h, int64_t N) {
    __shared__ double hs[2 * MAXV];
    if (threadIdx.x < 2 * nvec) hs[threadIdx.x] = h[threadIdx.x];
    __syncthreads();
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= N) return;
    double a0 = w0[n], a1 = w1[n];
    for (int i = 0; i < nvec; ++i) {
        const double v = V[(size_t)i * ldv + n];
        a0 -= hs[2 * i] * v;
        a1 -= hs[2 * i + 1] * v;
    }
    w0[n] = a0;
    w1[n] = a1;
}

// beta = sqrt(nrm2[0]); Hcol[j+1] = beta; v_next = w / beta; z_next = bw / beta (optional)
__global__ void k_normalize(const double* __restrict__ w, const double* __restrict__ bw, const double* __restrict__ nrm2,
                            double* __restrict__ hslot, double* __restrict__ vnext, double* __restrict__ znext,
                            int64_t N) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const double beta = sqrt(fabs(nrm2[0]));
    if (n == 0 && hslot) *hslot = beta;
    if (n >= N) return;
    const double inv = beta > 0.0 ? 1.0 / beta : 0.0;
    vnext[n] = w[n] * inv;
    if (znext) znext[n] = bw[n] * inv;
}

__global__ void k_random(double* __restrict__ v, int64_t N, uint64_t seed) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= N) return;
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(n + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    v[n] = ((double)(z >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;   // uniform(-1, 1) like ARPACK's v0
}

// in-place basis rotation (K10): rows 0..k-1 <- Q^T rows 0..m-1, per column of the basis;
// row k <- row m (the residual direction).  Q is m x k row-major in device memory.
//
// Wide bases (m > 32; Nmax = 80 / 100 of lantern-demo.ipynb gives m = 164 / 204): the rotation is a GEMM
// (k x m) (m x N) with 2 m k / (8 (m + k)) flop per byte -- 21 flop/B at m = 204, k = 150, above the fp64
// ridge of the B200 (~5.6 flop/B) -- so it runs on the fp64 tensor pipe (DMMA m8n8k4; tcgen05 has no fp64
// kind).  A CTA stages RT_TN columns of ALL m basis rows in shared memory (which is what makes the
// rotation safe in place: the tile is complete before its first row is overwritten), streams Q in slabs
// of RT_QC output rows and accumulates 2 x 2 MMA tiles per warp (8 warps = 2 x 4 over a 32 x 64 slab).
constexpr int RT_TN = 64, RT_VS = RT_TN + 4, RT_QC = 32, RT_QS = RT_QC + 4;
static size_t rotate_mma_smem(int m) { return (size_t)((m + 3) & ~3) * (RT_VS + RT_QS) * sizeof(double); }
__global__ void __launch_bounds__(256) k_rotate_mma(double* __restrict__ V, int64_t ldv, int m, int k,
                                                    const double* __restrict__ Q, int64_t N, int move_residual,
                                                    double* __restrict__ out, int64_t ldo) {
    extern __shared__ double rsm[];
    const int mp = (m + 3) & ~3;
    double* Vs = rsm;                          // [mp][RT_VS]  basis tile, row j = basis vector j
    double* Qs = rsm + (size_t)mp * RT_VS;     // [mp][RT_QS]  Q slab: Qs[j][ii] = Q[j][i0 + ii]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
    const int64_t n0 = (int64_t)blockIdx.x * RT_TN;
    const int len = (int)min((int64_t)RT_TN, N - n0);
    for (int j = warp; j < mp; j += 8) {
        const double* src = V + (size_t)j * ldv + n0;
#pragma unroll
        for (int c = lane; c < RT_TN; c += 32) Vs[j * RT_VS + c] = (j < m && c < len) ? src[c] : 0.0;
    }
    double res = 0.0;
    if (move_residual && tid < len) res = V[(size_t)m * ldv + n0 + tid];
    const int wm = warp & 1, wn = warp >> 1;
    double* dst = out ? out : V;
    const int64_t ldd = out ? ldo : ldv;
    for (int i0 = 0; i0 < k; i0 += RT_QC) {
        __syncthreads();                       // tile staged (first slab) / previous slab consumed
        for (int idx = tid; idx < mp * RT_QC; idx += 256) {
            const int j = idx >> 5, ii = idx & 31;
            Qs[j * RT_QS + ii] = (j < m && i0 + ii < k) ? Q[(size_t)j * k + i0 + ii] : 0.0;
        }
        __syncthreads();
        double acc[2][2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
        const double* qa = Qs + tq * RT_QS + wm * 16 + g;
        const double* vb = Vs + tq * RT_VS + wn * 16 + g;
#pragma unroll 4
        for (int j0 = 0; j0 < mp; j0 += 4, qa += 4 * RT_QS, vb += 4 * RT_VS) {
            const double a0 = qa[0], a1 = qa[8], b0 = vb[0], b1 = vb[8];
            dmma8x8x4(acc[0][0][0], acc[0][0][1], a0, b0);
            dmma8x8x4(acc[0][1][0], acc[0][1][1], a0, b1);
            dmma8x8x4(acc[1][0][0], acc[1][0][1], a1, b0);
            dmma8x8x4(acc[1][1][0], acc[1][1][1], a1, b1);
        }
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int row = i0 + wm * 16 + mb * 8 + g, col = wn * 16 + nb * 8 + 2 * tq + e;
                    if (row < k && col < len) dst[(size_t)row * ldd + n0 + col] = acc[mb][nb][e];
                }
    }
    if (move_residual && tid < len) V[(size_t)k * ldv + n0 + tid] = res;
}

// the same rotation with the basis column held in REGISTERS (MM >= m, fully unrolled, statically
// indexed): the generic kernel above indexes col[] dynamically, which puts it in local memory
template <int MM>
__global__ void __launch_bounds__(128) k_rotate_t(double* __restrict__ V, int64_t ldv, int m, int k,
                                                  const double* __restrict__ Q, int64_t N, int move_residual,
                                                  double* __restrict__ out, int64_t ldo) {
    extern __shared__ double qs[];
    for (int i = threadIdx.x; i < m * k; i += blockDim.x) qs[i] = Q[i];
    __syncthreads();
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= N) return;
    double col[MM];
#pragma unroll
    for (int j = 0; j < MM; ++j) col[j] = j < m ? V[(size_t)j * ldv + n] : 0.0;
    const double res = move_residual ? V[(size_t)m * ldv + n] : 0.0;
    double* dst = out ? out : V;
    const int64_t ldd = out ? ldo : ldv;
    for (int i = 0; i < k; ++i) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int j = 0; j < MM; j += 2) {
            if (j < m) s0 += qs[j * k + i] * col[j];
            if (j + 1 < m) s1 += qs[(j + 1) * k + i] * col[j + 1];
        }
        dst[(size_t)i * ldd + n] = s0 + s1;
    }
    if (move_residual) V[(size_t)k * ldv + n] = res;
}

static void launch_rotate(double* V, int64_t ldv, int m, int k, const double* Q, int64_t N, int move_residual, double* out,
                          int64_t ldo, cudaStream_t st) {
    const size_t sh = (size_t)m * k * sizeof(double);
    const int grid = ceil_div(N, 128);
    if (m <= 24) k_rotate_t<24><<<grid, 128, sh, st>>>(V, ldv, m, k, Q, N, move_residual, out, ldo);
    else if (m <= 32) k_rotate_t<32><<<grid, 128, sh, st>>>(V, ldv, m, k, Q, N, move_residual, out, ldo);
    else {
        static PerDeviceOnce configure;
        configure([] { WS_CUDA(cudaFuncSetAttribute(k_rotate_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rotate_mma_smem(MAX_NCV))); });
        k_rotate_mma<<<ceil_div(N, RT_TN), 256, rotate_mma_smem(m), st>>>(V, ldv, m, k, Q, N, move_residual, out, ldo);
    }
}

__global__ void k_scale_rows(double* __restrict__ X, int64_t ld, int nrows, const double* __restrict__ nrm2, int64_t N) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= N) return;
    for (int i = 0; i < nrows; ++i) {
        const double s = nrm2[i];
        X[(size_t)i * ld + n] *= (s > 0.0 ? 1.0 / sqrt(s) : 1.0);
    }
}

struct EigWork {
    int64_t N;
    int ncv;
    cudaStream_t st;
    DevBuf<double> V, Z, w, t1, t2, t3, partial, pnorm, h, Hd, nrm, Q;
    int nchunks;
    int chunk = DOT_CHUNK;      // rows per CTA of the dot kernels: 2048, or 512 for small problems (more CTAs than SMs)
};
// (wide bases get their parallelism from the vector groups of k_dots_partial and prefer the long chunk: Nmax = 80 at
//  N = 100k runs 493 ms with 2048-row chunks, 531 ms with 512)
static int dot_chunk_for(int64_t N, int m) { return (m <= 64 && N < 148 * 4 * (int64_t)DOT_CHUNK / 2) ? 512 : DOT_CHUNK; }

static void dots(EigWork& W, cudaStream_t q, const double* Zb, int nvec, const double* u, double* h, double* hacc,
                 FusedT tf = FusedT()) {
    int groups = 1;
    if (W.nchunks < 2 * kNumSMs && nvec > 8) groups = std::min((nvec + 7) / 8, (2 * kNumSMs + W.nchunks - 1) / W.nchunks);
    k_dots_partial<<<dim3(W.nchunks, groups), 256, 0, q>>>(Zb, W.N, nvec, u, W.N, W.partial.get(), W.chunk, tf);
    k_dots_reduce<<<nvec, 32, 0, q>>>(W.partial.get(), W.nchunks, nvec, h, hacc);
    count_launch(2);
}

}  // namespace ws

using namespace ws;

extern "C" {

// test hook: host dense eigen-solver (no GPU needed)
int ws_host_eig_general(int n, const double* h_A, double* h_wr, double* h_wi, double* h_V, int symmetric) {
    WS_API_BEGIN
    WS_REQUIRE(n >= 1 && n <= 512 && h_A && h_wr && h_V, "bad arguments");
    std::vector<double> A(h_A, h_A + (size_t)n * n), wr, wi, V;
    if (symmetric) {
        jacobi_sym(n, A, wr, V);
        wi.assign(n, 0.0);
    } else if (!eig_general(n, A, wr, wi, V)) {
        throw Error(WS_ERR_NOCONV, "dense QR iteration did not converge");
    }
    std::copy(wr.begin(), wr.end(), h_wr);
    if (h_wi) std::copy(wi.begin(), wi.end(), h_wi);
    std::copy(V.begin(), V.end(), h_V);
    WS_API_END
}

// ---------------------------------------------------------------------------------------------
// Block (size 2) Arnoldi with thick restart for the vector problem (kind 1).  The operator
// application is bound by the bytes of the inverted panels, which one sweep can share between two
// right-hand sides (ws_solver_solve, nrhs = 2): a block step applies OP to two basis vectors for
// little more than the price of one.  Basis column c is OP v_c orthogonalised against v_0..v_{c+1}
// and normalised into slot c+2, so H is band-Hessenberg with two subdiagonals and the residual of
// a Ritz pair is ||R y||, R = rows m, m+1 of the (m+2) x m projected matrix.
// ---------------------------------------------------------------------------------------------
static void eig_solve_block2(const ws_pattern* p, ws_solver* s, const double* B_vals, double sigma, int nev, int m,
                             double tol, int maxit, uint64_t seed, const int32_t* edges, const double* xy, int64_t Ntt,
                             double* w_out, double* V_out, int64_t* h_info, double* h_resid, int device, cudaStream_t st) {
    const int64_t N = p->N;
    const int LD = m + 2;
    EigWork W;
    W.N = N; W.ncv = m; W.st = st;
    W.nchunks = ceil_div(N, DOT_CHUNK);
    W.V.alloc((size_t)(m + 2) * N, st);
    W.w.alloc(2 * N, st); W.t1.alloc(2 * N, st); W.t2.alloc(2 * N, st); W.t3.alloc(2 * N, st);
    W.partial.alloc((size_t)W.nchunks * 2 * MAXV, st);
    W.h.alloc(2 * MAXV, st);
    W.Hd.alloc((size_t)LD * m, st);
    W.nrm.alloc(MAXV, st);
    W.Q.alloc((size_t)MAXV * MAXV, st);
    double* V = W.V.get();
    const int T = 256;
    const int nb = ceil_div(N, T);
    auto vecT = [&](cudaStream_t q, int mode, const double* in, double* out) {
        int rc = ws_vector_T(p, edges, xy, Ntt, mode, in, out, device, (void*)q);
        if (rc) throw Error(rc, ws_last_error());
    };
    // OP applied to basis vectors j, j+1 (contiguous rows of V) -> W.w (two vectors)
    auto apply_op2 = [&](cudaStream_t q, int j, int nrhs) {
        if (nrhs == 2) spmv_launch2(p, B_vals, V + (size_t)j * N, W.t1.get(), q);      // B read once for both
        else spmv_launch(p, B_vals, V + (size_t)j * N, W.t1.get(), q);
        for (int r = 0; r < nrhs; ++r) vecT(q, 1, W.t1.get() + (size_t)r * N, W.t2.get() + (size_t)r * N);
        int rc = ws_solver_solve(s, W.t2.get(), W.t3.get(), nrhs, device, (void*)q);
        if (rc) throw Error(rc, ws_last_error());
        for (int r = 0; r < nrhs; ++r) vecT(q, 0, W.t3.get() + (size_t)r * N, W.w.get() + (size_t)r * N);
    };
    // orthogonalise wv against rows 0..nvec-1 (CGS2), normalise into slot `slot`
    auto orth_into = [&](cudaStream_t q, double* wv, int nvec, int slot, double* Hcol) {
        for (int pass = 0; pass < 2 && nvec > 0; ++pass) {
            dots(W, q, V, nvec, wv, W.h.get(), Hcol);
            k_update<<<nb, T, (size_t)nvec * sizeof(double), q>>>(wv, V, N, nvec, W.h.get(), N);
            count_launch(1);
        }
        dots(W, q, wv, 1, wv, W.nrm.get(), nullptr);
        k_normalize<<<nb, T, 0, q>>>(wv, nullptr, W.nrm.get(), Hcol ? Hcol + slot : nullptr, V + (size_t)slot * N, nullptr, N);
        count_launch(1);
    };
    // block step: both new vectors against v_0..v_{j+1} in ONE pass over the basis per CGS pass, then the
    // second against the (normalised) first
    auto block_step = [&](cudaStream_t q, int j) {
        apply_op2(q, j, 2);
        double* w0 = W.w.get();
        double* w1 = W.w.get() + N;
        double* H0 = W.Hd.get() + (size_t)j * LD;
        double* H1 = W.Hd.get() + (size_t)(j + 1) * LD;
        const int nv = j + 2;
        for (int pass = 0; pass < 2; ++pass) {
            k_dots_partial2<<<W.nchunks, 256, 0, q>>>(V, N, nv, w0, w1, N, W.partial.get());
            k_dots_reduce2<<<2 * nv, 32, 0, q>>>(W.partial.get(), W.nchunks, nv, W.h.get(), H0, H1);
            k_update2<<<nb, T, 0, q>>>(w0, w1, V, N, nv, W.h.get(), N);
            count_launch(3);
        }
        dots(W, q, w0, 1, w0, W.nrm.get(), nullptr);
        k_normalize<<<nb, T, 0, q>>>(w0, nullptr, W.nrm.get(), H0 + (j + 2), V + (size_t)(j + 2) * N, nullptr, N);
        count_launch(1);
        for (int pass = 0; pass < 2; ++pass) {
            dots(W, q, V + (size_t)(j + 2) * N, 1, w1, W.h.get(), H1 + (j + 2));
            k_update<<<nb, T, sizeof(double), q>>>(w1, V + (size_t)(j + 2) * N, N, 1, W.h.get(), N);
            count_launch(1);
        }
        dots(W, q, w1, 1, w1, W.nrm.get(), nullptr);
        k_normalize<<<nb, T, 0, q>>>(w1, nullptr, W.nrm.get(), H1 + (j + 3), V + (size_t)(j + 3) * N, nullptr, N);
        count_launch(1);
    };
    struct GraphSet {
        std::vector<cudaGraphExec_t> exec;
        cudaStream_t cap = nullptr;
        ~GraphSet() {
            for (auto e : exec)
                if (e) cudaGraphExecDestroy(e);
            if (cap) cudaStreamDestroy(cap);
        }
    } graphs;
    graphs.exec.assign(m, nullptr);
    const char* nog = getenv("WS_NO_GRAPH");
    const bool use_graphs = !(nog && atoi(nog));
    if (use_graphs) WS_CUDA(cudaStreamCreateWithFlags(&graphs.cap, cudaStreamNonBlocking));
    auto run_step = [&](int j) {
        if (!use_graphs) {
            block_step(st, j);
            return;
        }
        if (!graphs.exec[j]) {
            cudaGraph_t g = nullptr;
            WS_CUDA(cudaStreamBeginCapture(graphs.cap, cudaStreamCaptureModeThreadLocal));
            try {
                block_step(graphs.cap, j);
            } catch (...) {
                cudaStreamEndCapture(graphs.cap, &g);
                if (g) cudaGraphDestroy(g);
                throw;
            }
            WS_CUDA(cudaStreamEndCapture(graphs.cap, &g));
            cudaError_t e = cudaGraphInstantiate(&graphs.exec[j], g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) throw Error(WS_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        }
        WS_CUDA(cudaGraphLaunch(graphs.exec[j], st));
        count_launch(1);
    };
    // ---- start block: two random vectors pushed once through OP (uncaptured: loads every kernel)
    k_random<<<nb, T, 0, st>>>(V, N, seed ? seed : 0x5DEECE66Dull);
    k_random<<<nb, T, 0, st>>>(V + N, N, (seed ? seed : 0x5DEECE66Dull) ^ 0x9E3779B97F4A7C15ull);
    count_launch(2);
    apply_op2(st, 0, 2);
    orth_into(st, W.w.get(), 0, 0, nullptr);
    orth_into(st, W.w.get() + N, 1, 1, nullptr);
    WS_CUDA(cudaMemsetAsync(W.Hd.get(), 0, W.Hd.bytes(), st));

    std::vector<double> Hh((size_t)LD * m, 0.0);
    std::vector<double> theta_r(m), theta_i(m), Y, resid(m, 0.0);
    std::vector<int> order(m);
    int j0 = 0, nconv = 0, restarts = 0, kkeep = 0;
    int64_t nops = 2;
    const double eps23 = std::pow(2.220446049250313e-16, 2.0 / 3.0);
    std::vector<double> Hm((size_t)m * m);
    while (true) {
        for (int j = j0; j < m; j += 2) {
            run_step(j);
            nops += 2;
        }
        WS_CUDA(cudaMemcpyAsync(Hh.data(), W.Hd.get(), Hh.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
        WS_CUDA(cudaStreamSynchronize(st));
        WS_CUDA(cudaGetLastError());
        bool finite = true;
        for (double v : Hh) finite = finite && std::isfinite(v);
        if (!finite) throw Error(WS_ERR_NUMERIC, "non-finite projected matrix in the block Krylov iteration");
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) Hm[(size_t)i * m + j] = Hh[(size_t)j * LD + i];
        std::vector<double> A = Hm;
        if (!eig_general(m, A, theta_r, theta_i, Y)) throw Error(WS_ERR_NOCONV, "dense QR iteration did not converge");
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return theta_r[a] < theta_r[b]; });
        // residual of a Ritz pair: || R y || / || y ||,  R = rows m, m+1 of the projected matrix
        auto rdot = [&](int row, int c) {
            double sacc = 0.0;
            for (int j = 0; j < m; ++j) sacc += Hh[(size_t)j * LD + row] * Y[(size_t)j * m + c];
            return sacc;
        };
        auto colnorm2 = [&](int c) {
            double sacc = 0.0;
            for (int i = 0; i < m; ++i) sacc += Y[(size_t)i * m + c] * Y[(size_t)i * m + c];
            return sacc;
        };
        auto ritz_res = [&](int c) {
            if (theta_i[c] == 0.0) return std::sqrt((rdot(m, c) * rdot(m, c) + rdot(m + 1, c) * rdot(m + 1, c)) / colnorm2(c));
            const int c0 = theta_i[c] > 0 ? c : c - 1;
            const double a0 = rdot(m, c0), a1 = rdot(m, c0 + 1), b0 = rdot(m + 1, c0), b1 = rdot(m + 1, c0 + 1);
            return std::sqrt((a0 * a0 + a1 * a1 + b0 * b0 + b1 * b1) / (colnorm2(c0) + colnorm2(c0 + 1)));
        };
        nconv = 0;
        for (int i = 0; i < nev; ++i) {
            const int c = order[i];
            resid[i] = ritz_res(c);
            const double mag = std::hypot(theta_r[c], theta_i[c]);
            if (resid[i] <= tol * std::max(eps23, mag)) ++nconv;
        }
        ++restarts;
        if (nconv >= nev || restarts >= maxit) break;
        // ---- thick restart: keep an even number of Ritz directions
        int k = nev + std::min(nconv, (m - nev) / 2);
        k += k & 1;
        if (k > m - 2) k = m - 2;
        std::vector<int> keep;
        {
            std::vector<char> used(m, 0);
            for (int i = 0; i < m && (int)keep.size() < k; ++i) {
                const int c = order[i];
                if (used[c]) continue;
                if (theta_i[c] == 0.0) {
                    keep.push_back(c);
                    used[c] = 1;
                } else {
                    const int c0 = theta_i[c] > 0.0 ? c : c - 1;
                    if ((int)keep.size() + 2 > m - 2) break;
                    keep.push_back(c0);
                    keep.push_back(c0 + 1);
                    used[c0] = used[c0 + 1] = 1;
                }
            }
            k = (int)keep.size();
        }
        std::vector<double> Qh((size_t)m * k, 0.0);
        for (int i = 0; i < k; ++i)
            for (int r = 0; r < m; ++r) Qh[(size_t)r * k + i] = Y[(size_t)r * m + keep[i]];
        int kk = 0;
        for (int i = 0; i < k; ++i) {
            for (int pass = 0; pass < 2; ++pass)
                for (int jq = 0; jq < kk; ++jq) {
                    double d = 0.0;
                    for (int r = 0; r < m; ++r) d += Qh[(size_t)r * k + jq] * Qh[(size_t)r * k + i];
                    for (int r = 0; r < m; ++r) Qh[(size_t)r * k + i] -= d * Qh[(size_t)r * k + jq];
                }
            double nn = 0.0;
            for (int r = 0; r < m; ++r) nn += Qh[(size_t)r * k + i] * Qh[(size_t)r * k + i];
            nn = std::sqrt(nn);
            if (nn < 1e-10) continue;
            for (int r = 0; r < m; ++r) Qh[(size_t)r * k + kk] = Qh[(size_t)r * k + i] / nn;
            ++kk;
        }
        kk -= kk & 1;                      // the remaining columns are processed in pairs
        if (kk < 2) throw Error(WS_ERR_NOCONV, "block Krylov restart lost its basis");
        std::vector<double> Qc((size_t)m * kk);
        for (int r = 0; r < m; ++r)
            for (int i = 0; i < kk; ++i) Qc[(size_t)r * kk + i] = Qh[(size_t)r * k + i];
        kkeep = kk;
        std::vector<double> HQ((size_t)m * kk, 0.0);
        for (int r = 0; r < m; ++r)
            for (int c = 0; c < kk; ++c) {
                double sacc = 0.0;
                for (int q2 = 0; q2 < m; ++q2) sacc += Hm[(size_t)r * m + q2] * Qc[(size_t)q2 * kk + c];
                HQ[(size_t)r * kk + c] = sacc;
            }
        std::vector<double> Hn((size_t)LD * m, 0.0);
        for (int c = 0; c < kk; ++c) {
            for (int r = 0; r < kk; ++r) {
                double sacc = 0.0;
                for (int q2 = 0; q2 < m; ++q2) sacc += Qc[(size_t)q2 * kk + r] * HQ[(size_t)q2 * kk + c];
                Hn[(size_t)c * LD + r] = sacc;
            }
            for (int rr = 0; rr < 2; ++rr) {
                double sacc = 0.0;
                for (int q2 = 0; q2 < m; ++q2) sacc += Hh[(size_t)q2 * LD + m + rr] * Qc[(size_t)q2 * kk + c];
                Hn[(size_t)c * LD + kk + rr] = sacc;
            }
        }
        Hh = Hn;
        WS_CUDA(cudaMemcpyAsync(W.Hd.get(), Hh.data(), Hh.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        WS_CUDA(cudaMemcpyAsync(W.Q.get(), Qc.data(), Qc.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        launch_rotate(V, N, m, kk, W.Q.get(), N, 1, nullptr, 0, st);                 // rows 0..kk-1, row m -> row kk
        WS_CUDA(cudaMemcpyAsync(V + (size_t)(kk + 1) * N, V + (size_t)(m + 1) * N, N * sizeof(double), cudaMemcpyDeviceToDevice, st));
        count_launch(1);
        WS_CUDA(cudaStreamSynchronize(st));     // Qc / Hh are host temporaries
        j0 = kk;
    }
    // ---- Ritz vectors, ordered by descending lambda = sigma + 1/theta
    std::vector<int> sel(order.begin(), order.begin() + nev);
    std::vector<double> lam(nev), lam_i(nev);
    for (int i = 0; i < nev; ++i) {
        const std::complex<double> th(theta_r[sel[i]], theta_i[sel[i]]);
        const std::complex<double> l = sigma + 1.0 / th;
        lam[i] = l.real();
        lam_i[i] = l.imag();
    }
    std::vector<int> ord2(nev);
    std::iota(ord2.begin(), ord2.end(), 0);
    std::stable_sort(ord2.begin(), ord2.end(), [&](int a, int b) { return lam[a] > lam[b]; });
    std::vector<double> Qx((size_t)m * nev), wh(nev), rh(nev);
    int ncomplex = 0;
    for (int i = 0; i < nev; ++i) {
        const int c = sel[ord2[i]];
        wh[i] = lam[ord2[i]];
        rh[i] = resid[ord2[i]];
        if (lam_i[ord2[i]] != 0.0) ++ncomplex;
        const int c0 = theta_i[c] < 0 ? c - 1 : c;
        double nn = 0.0;
        for (int r = 0; r < m; ++r) nn += Y[(size_t)r * m + c0] * Y[(size_t)r * m + c0];
        nn = std::sqrt(nn);
        for (int r = 0; r < m; ++r) Qx[(size_t)r * nev + i] = Y[(size_t)r * m + c0] / nn;
    }
    WS_CUDA(cudaMemcpyAsync(W.Q.get(), Qx.data(), Qx.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    launch_rotate(V, N, m, nev, W.Q.get(), N, 0, V_out, N, st);
    count_launch(1);
    for (int i = 0; i < nev; ++i) dots(W, st, V_out + (size_t)i * N, 1, V_out + (size_t)i * N, W.nrm.get() + i, nullptr);
    k_scale_rows<<<nb, T, 0, st>>>(V_out, N, nev, W.nrm.get(), N);
    count_launch(1);
    WS_CUDA(cudaMemcpyAsync(w_out, wh.data(), nev * sizeof(double), cudaMemcpyHostToDevice, st));
    WS_CUDA(cudaStreamSynchronize(st));
    WS_CUDA(cudaGetLastError());
    {
        int rc = ws_solver_check(s, device, (void*)st);
        if (rc) throw Error(rc, ws_last_error());
    }
    if (getenv("WS_TRACE"))
        fprintf(stderr, "[ws_eig block2] N=%lld ops=%lld (%lld sweeps) restarts=%d nconv=%d\n", (long long)N, (long long)nops,
                (long long)(nops / 2), restarts, nconv);
    if (h_info) {
        h_info[0] = nconv; h_info[1] = nops; h_info[2] = restarts; h_info[3] = ncomplex; h_info[4] = kkeep;
    }
    if (h_resid) std::copy(rh.begin(), rh.end(), h_resid);
    if (nconv < nev) throw Error(WS_ERR_NOCONV, "eigensolver: " + std::to_string(nconv) + " of " + std::to_string(nev) +
                                                    " eigenpairs converged in " + std::to_string(restarts) + " restarts");
}

// ---------------------------------------------------------------------------------------------
// The single-vector iteration as a RESUMABLE object.  One restart cycle = (enqueue the Arnoldi steps and
// the copy of the projected matrix) / (wait, analyse on the host, enqueue the restart rotation).  Nothing
// between the two halves blocks the host, so one host thread can drive several eigenproblems on several
// streams in lockstep (ws_eig_solve_batch): the kernels of a ~50k-triangle problem are a chain of small
// dependent launches that fills a fraction of a B200, and independent problems on other streams fill the
// rest (measured: 4 streams, 3.2 solves in the time of one).  ws_eig_solve is the same object driven alone.
// ---------------------------------------------------------------------------------------------
namespace {

struct PinnedBuf {
    double* p = nullptr;
    size_t n = 0;
    void alloc(size_t n_) {
        if (p && n == n_) return;           // (page-locking is slow: a work space that moves to the next mesh keeps its blocks)
        release();
        n = n_;
        if (n) WS_CUDA(cudaMallocHost((void**)&p, n * sizeof(double)));
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        n = 0;
    }
    ~PinnedBuf() { release(); }
};

struct GraphSet {
    std::vector<cudaGraphExec_t> exec;
    cudaStream_t cap = nullptr;
    ~GraphSet() {
        for (auto e : exec)
            if (e) cudaGraphExecDestroy(e);
        if (cap) cudaStreamDestroy(cap);
    }
};

struct EigRun {
    // problem
    const ws_pattern* p = nullptr;
    ws_solver* s = nullptr;
    const double* B_vals = nullptr;
    double sigma = 0.0;
    int kind = 0, nev = 0, m = 0, maxit = 0, device = 0;
    double tol = 0.0;
    uint64_t seed = 0;
    const int32_t* edges = nullptr;
    const double* xy = nullptr;
    int64_t Ntt = 0, N = 0;
    double* w_out = nullptr;
    double* V_out = nullptr;
    cudaStream_t st = nullptr;
    // device work space
    EigWork W;
    GraphSet graphs;
    bool use_graphs = true;
    double* V = nullptr;
    double* Z = nullptr;
    int nb = 0;
    static constexpr int T = 256;
    // host state
    PinnedBuf hH, hUp;                      // projected matrix coming back / (H, Q) going up
    cudaEvent_t ev = nullptr;
    std::vector<double> Hm, theta_r, theta_i, Y, resid;
    std::vector<int> order;
    int j0 = 0, nconv = 0, restarts = 0, kkeep = 0, ncomplex = 0;
    int64_t nops = 1;
    std::vector<double> wh, rh, lam_i_sorted;
    // tracing (WS_TRACE=1: synchronises after every phase, diagnostic only; WS_TRACE=2: events around every step)
    bool trace = false, etrace = false;
    std::vector<cudaEvent_t> evs;
    double wall0 = 0, tr_op = 0, tr_host = 0, tr_rot = 0, tr_t0 = 0;
    double t_capture = 0;                   // host time spent capturing / instantiating step graphs (this work space's life)
    int n_capture = 0;

    ~EigRun() {
        if (ev) cudaEventDestroy(ev);
        for (auto e : evs) cudaEventDestroy(e);
    }
    static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    double tick() {
        if (!trace) return 0.0;
        cudaStreamSynchronize(st);
        return now();
    }
    void mark() {
        if (!etrace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        evs.push_back(e);
    }

    void solve(cudaStream_t q, const double* b, double* x) {
        int rc = ws_solver_solve(s, b, x, 1, device, (void*)q);
        if (rc) throw Error(rc, ws_last_error());
    }
    void vecT(cudaStream_t q, int mode, const double* in, double* out) {
        int rc = ws_vector_T(p, edges, xy, Ntt, mode, in, out, device, (void*)q);
        if (rc) throw Error(rc, ws_last_error());
    }
    // OP applied to an arbitrary vector x (kind 0: bx = B x must be given) -> W.w
    void apply_op_vec(cudaStream_t q, const double* x, const double* bx) {
        if (kind == 0) {
            solve(q, bx, W.w.get());
        } else {
            spmv_launch(p, B_vals, x, W.t1.get(), q);
            vecT(q, 1, W.t1.get(), W.t2.get());
            solve(q, W.t2.get(), W.t3.get());
            vecT(q, 0, W.t3.get(), W.w.get());
        }
    }
    // OP applied to basis vector j -> W.w   (kind 0: Z_j = B v_j is carried along)
    void apply_op(cudaStream_t q, int j) { apply_op_vec(q, V + (size_t)j * N, kind == 0 ? Z + (size_t)j * N : nullptr); }
    // normalise W.w in the inner product into basis slot j (H entry written to hslot)
    void normalize_into(cudaStream_t q, int j, double* hslot) {
        const double* dual = W.w.get();
        if (kind == 0) {
            spmv_launch(p, B_vals, W.w.get(), W.t1.get(), q);
            dual = W.t1.get();
        }
        dots(W, q, dual, 1, W.w.get(), W.nrm.get(), nullptr);
        k_normalize<<<nb, T, 0, q>>>(W.w.get(), kind == 0 ? W.t1.get() : nullptr, W.nrm.get(), hslot, V + (size_t)j * N,
                                    kind == 0 ? Z + (size_t)j * N : nullptr, N);
        count_launch(1);
    }
    // one Arnoldi step: w = OP v_j, CGS2 against rows 0..j (coefficients with the dual basis, update with the
    // primal one), normalise into slot j+1
    void arnoldi_step(cudaStream_t q, int j) {
        static const bool fuse_cgs = !(getenv("WS_NO_CGS_FUSE") && atoi(getenv("WS_NO_CGS_FUSE")));
        static const bool fuse_glue = !(getenv("WS_NO_GLUE_FUSE") && atoi(getenv("WS_NO_GLUE_FUSE")));
        double* Hcol = W.Hd.get() + (size_t)j * (m + 1);
        const int nv = j + 1;
        if (kind == 1 && fuse_cgs && fuse_glue && nv + 1 <= MAXV) {
            // Fully fused vector step (7 launches around the solve instead of 12):
            //   B v_j -> T^T -> solve  |  [T fused into the first products]  h1 = V^T w
            //   |  w' = w - V h1, h2 = V^T w', ||w'||^2 in one pass over the basis
            //   |  beta^2 = ||w'||^2 - ||h2||^2  |  v_{j+1} = (w' - V h2) / beta in one pass
            spmv_launch(p, B_vals, V + (size_t)j * N, W.t1.get(), q);
            vecT(q, 1, W.t1.get(), W.t2.get());
            solve(q, W.t2.get(), W.t3.get());
            FusedT tf;
            tf.tin = W.t3.get();
            tf.edges = (const int2*)edges;
            tf.elen = p->vt_len.get();
            tf.Ntt = Ntt;
            tf.wout = W.w.get();
            dots(W, q, V, nv, W.w.get(), W.h.get(), Hcol, tf);
            k_update_dots<<<W.nchunks, 256, (size_t)(nv + 1) * UD_TILE * sizeof(double), q>>>(W.w.get(), V, N, nv, W.h.get(), N,
                                                                                         W.partial.get(), W.chunk, W.pnorm.get());
            k_dots_reduce_beta<<<1, 256, 0, q>>>(W.partial.get(), W.pnorm.get(), W.nchunks, nv, W.h.get(), Hcol, W.nrm.get());
            k_update_normalize<<<nb, T, (size_t)nv * sizeof(double), q>>>(W.w.get(), V, N, nv, W.h.get(), N, W.nrm.get(),
                                                                        Hcol + (j + 1), V + (size_t)(j + 1) * N);
            count_launch(3);
            return;
        }
        apply_op(q, j);
        if (kind == 1 && fuse_cgs && nv + 1 <= MAXV) {
            // Euclidean inner product (dual basis = basis): the update of pass 1 and the inner products of pass 2
            // share one pass over the basis -- three basis reads per step instead of four
            dots(W, q, V, nv, W.w.get(), W.h.get(), Hcol);
            k_update_dots<<<W.nchunks, 256, (size_t)(nv + 1) * UD_TILE * sizeof(double), q>>>(W.w.get(), V, N, nv, W.h.get(), N,
                                                                                         W.partial.get(), W.chunk, nullptr);
            k_dots_reduce<<<nv, 32, 0, q>>>(W.partial.get(), W.nchunks, nv, W.h.get(), Hcol);
            k_update<<<nb, T, (size_t)nv * sizeof(double), q>>>(W.w.get(), V, N, nv, W.h.get(), N);
            count_launch(3);
        } else {
            for (int pass = 0; pass < 2; ++pass) {
                dots(W, q, Z, nv, W.w.get(), W.h.get(), Hcol);
                k_update<<<nb, T, (size_t)nv * sizeof(double), q>>>(W.w.get(), V, N, nv, W.h.get(), N);
                count_launch(1);
            }
        }
        normalize_into(q, j + 1, Hcol + (j + 1));
    }
    // every enqueue helper takes the stream explicitly: the Arnoldi steps are recorded once into CUDA graphs
    // (stream capture on a private stream) and replayed on the caller's stream, so an eigensolve costs ~10^2
    // host launches instead of ~7000 and does not depend on host latency
    void run_step(int j) {
        // Steps below nev only ever run in the first restart cycle of a problem (a thick restart keeps at least nev
        // directions), so a graph of such a step is replayed once per problem: capturing + instantiating it (~0.4 ms
        // of host time) costs more than launching its ~36 kernels directly (~0.08 ms).  Graphs pay for the steps
        // that recur in every cycle.
        static const int graph_from = getenv("WS_GRAPH_FROM") ? atoi(getenv("WS_GRAPH_FROM")) : -1;
        if (!use_graphs || j < (graph_from >= 0 ? graph_from : nev)) {
            arnoldi_step(st, j);
            return;
        }
        if (!graphs.exec[j]) {
            const double tc0 = now();
            cudaGraph_t g = nullptr;
            WS_CUDA(cudaStreamBeginCapture(graphs.cap, cudaStreamCaptureModeThreadLocal));
            try {
                arnoldi_step(graphs.cap, j);
            } catch (...) {
                cudaStreamEndCapture(graphs.cap, &g);
                if (g) cudaGraphDestroy(g);
                throw;
            }
            WS_CUDA(cudaStreamEndCapture(graphs.cap, &g));
            cudaError_t e = cudaGraphInstantiate(&graphs.exec[j], g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) throw Error(WS_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
            t_capture += now() - tc0;
            ++n_capture;
        }
        WS_CUDA(cudaGraphLaunch(graphs.exec[j], st));
        count_launch(1);
    }

    // What the captured step graphs and the work space depend on.  A run object kept in a ws_eig_workspace is
    // reused as it is while this key stays the same (the wavelengths of one mesh on one sweep lane: same pattern,
    // solver, value array, sizes), so the ~ncv graph captures / instantiations, the pinned buffers and the basis are
    // paid once per mesh instead of once per problem.
    struct Key {
        const void *p = nullptr, *s = nullptr, *B = nullptr, *edges = nullptr, *xy = nullptr, *st = nullptr;
        int64_t N = 0, Ntt = 0, solver_uid = -1;
        int kind = -1, m = 0, refine = -1;
        bool operator==(const Key& o) const {
            return p == o.p && s == o.s && B == o.B && edges == o.edges && xy == o.xy && st == o.st && N == o.N && Ntt == o.Ntt &&
                   solver_uid == o.solver_uid && kind == o.kind && m == o.m && refine == o.refine;
        }
    } key;

    // allocate (or keep) the work space, then build the start vector (random, pushed once through OP: stays in
    // range(OP), as ARPACK does)
    void init() {
        N = p->N;
        int64_t qi[4] = {0, 0, 0, 0};
        ws_solver_quality(s, qi, nullptr);
        Key k2;
        k2.p = p; k2.s = s; k2.B = B_vals; k2.edges = edges; k2.xy = xy; k2.st = st; k2.N = N; k2.Ntt = Ntt; k2.kind = kind; k2.m = m;
        k2.refine = (int)qi[1];
        k2.solver_uid = qi[3];
        if (!(k2 == key)) {
            for (auto& e : graphs.exec)
                if (e) cudaGraphExecDestroy(e);
            W.N = N; W.ncv = m; W.st = st;
            W.chunk = dot_chunk_for(N, m);
            W.nchunks = ceil_div(N, W.chunk);
            W.V.alloc((size_t)(m + 1) * N, st);
            if (kind == 0) W.Z.alloc((size_t)(m + 1) * N, st);
            else W.Z.release();
            W.w.alloc(N, st); W.t1.alloc(N, st); W.t2.alloc(N, st); W.t3.alloc(N, st);
            W.partial.alloc((size_t)W.nchunks * (m + 1), st);
            W.pnorm.alloc((size_t)W.nchunks, st);
            W.h.alloc(m + 1, st);
            W.Hd.alloc((size_t)(m + 1) * m, st);
            W.nrm.alloc(m + 1, st);
            W.Q.alloc((size_t)m * m, st);
            hH.alloc((size_t)(m + 1) * m);
            hUp.alloc((size_t)(m + 1) * m + (size_t)m * m);
            if (!ev) WS_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            graphs.exec.assign(m, nullptr);
            key = k2;
        }
        V = W.V.get();
        Z = kind == 0 ? W.Z.get() : V;
        nb = ceil_div(N, T);
        j0 = nconv = restarts = kkeep = ncomplex = 0;
        nops = 1;
        tr_op = tr_host = tr_rot = 0;
        for (auto e : evs) cudaEventDestroy(e);
        evs.clear();
        const char* nog = getenv("WS_NO_GRAPH");
        use_graphs = !(nog && atoi(nog));
        if (use_graphs && !graphs.cap) WS_CUDA(cudaStreamCreateWithFlags(&graphs.cap, cudaStreamNonBlocking));
        const char* trace_env = getenv("WS_TRACE");
        trace = trace_env != nullptr && atoi(trace_env) == 1;
        etrace = trace_env != nullptr && atoi(trace_env) == 2;
        wall0 = now();
        tr_t0 = tick();
        k_random<<<nb, T, 0, st>>>(W.w.get(), N, seed ? seed : 0x5DEECE66Dull);
        count_launch(1);
        normalize_into(st, 0, nullptr);
        apply_op(st, 0);
        normalize_into(st, 0, nullptr);
        // every kernel of a step has now run once outside stream capture, except the update:
        k_update<<<nb, T, 0, st>>>(W.w.get(), V, N, 0, W.h.get(), N);   // nvec = 0: no-op, loads the kernel
        if (kind == 1) {
            // the same for the kernels of the fused step (harmless with nvec = 0: w unchanged, scratch outputs)
            k_update_dots<<<W.nchunks, 256, (size_t)UD_TILE * sizeof(double), st>>>(W.w.get(), V, N, 0, W.h.get(), N, W.partial.get(),
                                                                                 W.chunk, W.pnorm.get());
            k_dots_reduce_beta<<<1, 256, 0, st>>>(W.partial.get(), W.pnorm.get(), W.nchunks, 0, W.h.get(), nullptr, W.nrm.get());
            k_update_normalize<<<nb, T, 0, st>>>(W.w.get(), V, N, 0, W.h.get(), N, W.nrm.get(), nullptr, W.t1.get());
            count_launch(3);
        }
        WS_CUDA(cudaMemsetAsync(W.Hd.get(), 0, W.Hd.bytes(), st));
        theta_r.assign(m, 0.0);
        theta_i.assign(m, 0.0);
        resid.assign(m, 0.0);
        order.assign(m, 0);
        Hm.assign((size_t)m * m, 0.0);
    }

    // first half of a restart cycle: steps j0 .. m-1, projected matrix -> pinned host memory; never blocks
    void enqueue_cycle() {
        for (int j = j0; j < m; ++j) {
            double ta = tick();
            mark();
            run_step(j);
            mark();
            ++nops;
            tr_op += tick() - ta;
        }
        WS_CUDA(cudaMemcpyAsync(hH.p, W.Hd.get(), (size_t)(m + 1) * m * sizeof(double), cudaMemcpyDeviceToHost, st));
        WS_CUDA(cudaEventRecord(ev, st));
    }

    // second half: wait for the cycle, Ritz analysis; returns true when the iteration is over, otherwise the thick
    // restart (host) and its basis rotation (enqueued) have been set up for the next cycle
    bool finish_cycle() {
        double th0 = tick();
        WS_CUDA(cudaEventSynchronize(ev));
        WS_CUDA(cudaGetLastError());
        const double* Hh = hH.p;                         // column-major (m+1) x m, as on the device
        const double beta_m = Hh[(size_t)(m - 1) * (m + 1) + m];
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) Hm[(size_t)i * m + j] = Hh[(size_t)j * (m + 1) + i];
        bool finite = std::isfinite(beta_m);
        for (double v : Hm) finite = finite && std::isfinite(v);
        if (!finite) throw Error(WS_ERR_NUMERIC, "non-finite projected matrix in the Krylov iteration");
        if (kind == 0) {
            for (int i = 0; i < m; ++i)
                for (int j = i + 1; j < m; ++j) {
                    const double a = 0.5 * (Hm[(size_t)i * m + j] + Hm[(size_t)j * m + i]);
                    Hm[(size_t)i * m + j] = Hm[(size_t)j * m + i] = a;
                }
            std::vector<double> A = Hm;
            jacobi_sym(m, A, theta_r, Y);
            std::fill(theta_i.begin(), theta_i.end(), 0.0);
        } else {
            std::vector<double> A = Hm;
            if (!eig_general(m, A, theta_r, theta_i, Y)) throw Error(WS_ERR_NOCONV, "dense QR iteration did not converge");
        }
        // wanted: smallest real part of theta = 1/(lambda - sigma)  ('SA' / 'SR')
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return theta_r[a] < theta_r[b]; });
        // Ritz residual estimates |beta_m * (last component of the unit eigenvector)|
        auto colnorm = [&](int c, bool cplx) {
            double sacc = 0.0;
            for (int i = 0; i < m; ++i) {
                sacc += Y[(size_t)i * m + c] * Y[(size_t)i * m + c];
                if (cplx) sacc += Y[(size_t)i * m + c + 1] * Y[(size_t)i * m + c + 1];
            }
            return std::sqrt(sacc);
        };
        auto ritz_res = [&](int c) {
            if (theta_i[c] == 0.0) return std::fabs(beta_m * Y[(size_t)(m - 1) * m + c]) / colnorm(c, false);
            const int c0 = theta_i[c] > 0 ? c : c - 1;
            const double lr = Y[(size_t)(m - 1) * m + c0], li = Y[(size_t)(m - 1) * m + c0 + 1];
            return std::fabs(beta_m) * std::sqrt(lr * lr + li * li) / colnorm(c0, true);
        };
        const double eps23 = std::pow(2.220446049250313e-16, 2.0 / 3.0);
        nconv = 0;
        for (int i = 0; i < nev; ++i) {
            const int c = order[i];
            resid[i] = ritz_res(c);
            const double mag = std::hypot(theta_r[c], theta_i[c]);
            if (resid[i] <= tol * std::max(eps23, mag)) ++nconv;
        }
        ++restarts;
        if (nconv >= nev || restarts >= maxit) return true;
        // ---- thick restart: keep k Ritz directions (ARPACK-style: more as they converge)
        static const int keep_policy = getenv("WS_EIG_KEEP") ? atoi(getenv("WS_EIG_KEEP")) : 0;
        int k = nev + std::min(nconv, (m - nev) / 2);
        if (keep_policy == 1) k = nev + (m - nev) / 2;
        else if (keep_policy == 2) k = nev + (m - nev) / 3;
        else if (keep_policy == 3) k = nev + std::max(std::min(nconv, (m - nev) / 2), (m - nev) / 4);
        if (k >= m - 1) k = m - 2;
        // kept columns of Y; a complex-conjugate pair contributes its real and imaginary parts
        std::vector<int> keep;
        {
            std::vector<char> used(m, 0);
            for (int i = 0; i < m && (int)keep.size() < k; ++i) {
                const int c = order[i];
                if (used[c]) continue;
                if (theta_i[c] == 0.0) {
                    keep.push_back(c);
                    used[c] = 1;
                } else {
                    const int c0 = theta_i[c] > 0.0 ? c : c - 1;
                    if ((int)keep.size() + 2 > m - 2) break;
                    keep.push_back(c0);
                    keep.push_back(c0 + 1);
                    used[c0] = used[c0 + 1] = 1;
                }
            }
            k = (int)keep.size();
        }
        // orthonormal basis Q (m x k) of the kept invariant subspace (modified Gram-Schmidt, twice)
        std::vector<double> Qh((size_t)m * k, 0.0);
        for (int i = 0; i < k; ++i) {
            const int c = keep[i];
            for (int r = 0; r < m; ++r) Qh[(size_t)r * k + i] = Y[(size_t)r * m + c];
        }
        int kk = 0;
        for (int i = 0; i < k; ++i) {
            for (int pass = 0; pass < 2; ++pass)
                for (int jq = 0; jq < kk; ++jq) {
                    double d = 0.0;
                    for (int r = 0; r < m; ++r) d += Qh[(size_t)r * k + jq] * Qh[(size_t)r * k + i];
                    for (int r = 0; r < m; ++r) Qh[(size_t)r * k + i] -= d * Qh[(size_t)r * k + jq];
                }
            double nn = 0.0;
            for (int r = 0; r < m; ++r) nn += Qh[(size_t)r * k + i] * Qh[(size_t)r * k + i];
            nn = std::sqrt(nn);
            if (nn < 1e-10) continue;     // linearly dependent direction: drop it
            for (int r = 0; r < m; ++r) Qh[(size_t)r * k + kk] = Qh[(size_t)r * k + i] / nn;
            ++kk;
        }
        kkeep = kk;
        // (H, Q) of the restarted factorisation go up through pinned memory: [Q^T H Q ; beta_m * (last row of Q)]
        double* Hup = hUp.p;
        double* Qc = hUp.p + (size_t)(m + 1) * m;        // m x kk row-major
        for (int r = 0; r < m; ++r)
            for (int i = 0; i < kk; ++i) Qc[(size_t)r * kk + i] = Qh[(size_t)r * k + i];
        std::vector<double> HQ((size_t)m * kk, 0.0);
        for (int r = 0; r < m; ++r)
            for (int q2 = 0; q2 < m; ++q2) {
                const double hv = Hm[(size_t)r * m + q2];
                if (hv == 0.0) continue;
                const double* qrow = Qc + (size_t)q2 * kk;
                double* dst = HQ.data() + (size_t)r * kk;
                for (int c = 0; c < kk; ++c) dst[c] += hv * qrow[c];
            }
        std::fill(Hup, Hup + (size_t)(m + 1) * m, 0.0);
        for (int c = 0; c < kk; ++c) {
            for (int r = 0; r < kk; ++r) {
                double sacc = 0.0;
                for (int q2 = 0; q2 < m; ++q2) sacc += Qc[(size_t)q2 * kk + r] * HQ[(size_t)q2 * kk + c];
                Hup[(size_t)c * (m + 1) + r] = sacc;
            }
            Hup[(size_t)c * (m + 1) + kk] = beta_m * Qc[(size_t)(m - 1) * kk + c];
        }
        double th1 = tick();
        tr_host += th1 - th0;
        WS_CUDA(cudaMemcpyAsync(W.Hd.get(), Hup, (size_t)(m + 1) * m * sizeof(double), cudaMemcpyHostToDevice, st));
        WS_CUDA(cudaMemcpyAsync(W.Q.get(), Qc, (size_t)m * kk * sizeof(double), cudaMemcpyHostToDevice, st));
        launch_rotate(V, N, m, kk, W.Q.get(), N, 1, nullptr, 0, st);
        if (kind == 0) launch_rotate(Z, N, m, kk, W.Q.get(), N, 1, nullptr, 0, st);
        count_launch(kind == 0 ? 2 : 1);
        // (hUp is rewritten only after the next cycle's event: that cycle is enqueued behind these copies)
        tr_rot += tick() - th1;
        j0 = kk;
        return false;
    }

    // Ritz vectors of the nev wanted eigenvalues, ordered by descending lambda = sigma + 1/theta; enqueued, no sync
    void enqueue_results() {
        std::vector<int> sel(order.begin(), order.begin() + nev);
        std::vector<double> lam(nev), lam_i(nev);
        for (int i = 0; i < nev; ++i) {
            const std::complex<double> th(theta_r[sel[i]], theta_i[sel[i]]);
            const std::complex<double> l = sigma + 1.0 / th;
            lam[i] = l.real();
            lam_i[i] = l.imag();
        }
        std::vector<int> ord2(nev);
        std::iota(ord2.begin(), ord2.end(), 0);
        std::stable_sort(ord2.begin(), ord2.end(), [&](int a, int b) { return lam[a] > lam[b]; });
        double* Qx = hUp.p + (size_t)(m + 1) * m;        // m x nev row-major
        double* whp = hUp.p;                              // nev eigenvalues
        wh.assign(nev, 0.0);
        rh.assign(nev, 0.0);
        lam_i_sorted.assign(nev, 0.0);
        ncomplex = 0;
        for (int i = 0; i < nev; ++i) {
            const int c = sel[ord2[i]];
            wh[i] = whp[i] = lam[ord2[i]];
            rh[i] = resid[ord2[i]];
            lam_i_sorted[i] = lam_i[ord2[i]];
            if (lam_i[ord2[i]] != 0.0) ++ncomplex;
            const int c0 = theta_i[c] < 0 ? c - 1 : c;    // complex pair: return the real part
            double nn = 0.0;
            for (int r = 0; r < m; ++r) nn += Y[(size_t)r * m + c0] * Y[(size_t)r * m + c0];
            nn = std::sqrt(nn);
            for (int r = 0; r < m; ++r) Qx[(size_t)r * nev + i] = Y[(size_t)r * m + c0] / nn;
        }
        WS_CUDA(cudaMemcpyAsync(W.Q.get(), Qx, (size_t)m * nev * sizeof(double), cudaMemcpyHostToDevice, st));
        launch_rotate(V, N, m, nev, W.Q.get(), N, 0, V_out, N, st);
        count_launch(1);
        if (kind == 1) {
            // unit 2-norm rows (what ARPACK's dneupd returns)
            for (int i = 0; i < nev; ++i) dots(W, st, V_out + (size_t)i * N, 1, V_out + (size_t)i * N, W.nrm.get() + i, nullptr);
            k_scale_rows<<<nb, T, 0, st>>>(V_out, N, nev, W.nrm.get(), N);
            count_launch(1);
        }
        WS_CUDA(cudaMemcpyAsync(w_out, whp, nev * sizeof(double), cudaMemcpyHostToDevice, st));
    }

    // wait for the results, solver sanity, true-residual check; fills the caller's diagnostics; throws on failure
    void finish(int64_t* h_info, double* h_resid) {
        WS_CUDA(cudaStreamSynchronize(st));
        WS_CUDA(cudaGetLastError());
        {
            int rc = ws_solver_check(s, device, (void*)st);
            if (rc) throw Error(rc, ws_last_error());
        }
        if (h_info) {
            h_info[0] = nconv; h_info[1] = nops; h_info[2] = restarts; h_info[3] = ncomplex; h_info[4] = kkeep;
        }
        if (h_resid) std::copy(rh.begin(), rh.end(), h_resid);
        // ---- true residuals of the returned pairs, || OP x - theta x || / (|theta| ||x||), with OP applied through
        // the (refined) solves.  The Ritz estimates are statements about the operator the iteration saw; when the
        // factorisation needed static pivots or refinement (interior shifts) the returned pairs are re-checked
        // against a fresh application, and the call fails rather than return pairs that miss 1e-9.
        // WS_EIG_VERIFY=1 / 0 forces / disables the check.
        int64_t qi[4] = {0, 0, 0, 0};
        double qd[4];
        ws_solver_quality(s, qi, qd);
        static const int verify_env = getenv("WS_EIG_VERIFY") ? atoi(getenv("WS_EIG_VERIFY")) : -1;
        const bool verify = verify_env >= 0 ? verify_env != 0 : (qi[0] > 0 || qi[1] > 0);
        if (verify && nconv >= nev) {
            DevBuf<double> vr((size_t)3 * nev, st);
            std::vector<double> th(nev), back((size_t)3 * nev, 0.0);
            for (int i = 0; i < nev; ++i) th[i] = 1.0 / (wh[i] - sigma);
            WS_CUDA(cudaMemcpyAsync(vr.get(), th.data(), nev * sizeof(double), cudaMemcpyHostToDevice, st));
            for (int i = 0; i < nev; ++i) {
                const double* xi = V_out + (size_t)i * N;
                const double* bx = nullptr;
                if (kind == 0) {
                    spmv_launch(p, B_vals, xi, W.t1.get(), st);
                    bx = W.t1.get();
                }
                apply_op_vec(st, xi, bx);
                k_update<<<nb, T, sizeof(double), st>>>(W.w.get(), xi, N, 1, vr.get() + i, N);
                count_launch(1);
                dots(W, st, W.w.get(), 1, W.w.get(), vr.get() + nev + i, nullptr);
                dots(W, st, xi, 1, xi, vr.get() + 2 * nev + i, nullptr);
            }
            WS_CUDA(cudaMemcpyAsync(back.data(), vr.get(), back.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
            WS_CUDA(cudaStreamSynchronize(st));
            double worst = 0.0;
            int worst_i = -1;
            for (int i = 0; i < nev; ++i) {
                if (lam_i_sorted[i] != 0.0) continue;           // complex pair: only the real part is returned
                const double r = std::sqrt(back[nev + i] / back[2 * nev + i]) / std::fabs(th[i]);
                if (!(r <= worst)) { worst = r; worst_i = i; }
            }
            if (getenv("WS_TRACE"))
                fprintf(stderr, "[ws_eig verify] worst true residual %.3e (pair %d), solver: %lld perturbed pivots, %lld refinement steps\n",
                        worst, worst_i, (long long)qi[0], (long long)qi[1]);
            if (!(worst <= 1e-9))
                throw Error(WS_ERR_NUMERIC, "eigenpair " + std::to_string(worst_i) + " fails the residual check against the operator (" +
                                                std::to_string(worst) + "): the factorisation at this shift is not accurate enough");
        }
        if (etrace) {
            const double wall1 = now();
            double op_ms = 0, mx = 0, mn = 1e30, span = 0;
            for (size_t i = 0; i + 1 < evs.size(); i += 2) {
                float ms = 0;
                cudaEventElapsedTime(&ms, evs[i], evs[i + 1]);
                op_ms += ms; mx = std::max<double>(mx, ms); mn = std::min<double>(mn, ms);
            }
            if (evs.size() >= 2) { float ms = 0; cudaEventElapsedTime(&ms, evs.front(), evs.back()); span = ms; }
            fprintf(stderr, "[ws_eig etrace] wall %.1f ms, first-op-start..last-op-end %.1f ms, sum(op) %.1f ms over %zu ops (min %.3f max %.3f)\n",
                    (wall1 - wall0) * 1e3, span, op_ms, evs.size() / 2, mn, mx);
        }
        if (trace)
            fprintf(stderr, "[ws_eig] kind=%d N=%lld ops=%lld restarts=%d total=%.1f ms: op=%.1f host=%.1f rotate=%.1f\n",
                    kind, (long long)N, (long long)nops, restarts, (tick() - tr_t0) * 1e3, tr_op * 1e3, tr_host * 1e3, tr_rot * 1e3);
        if (nconv < nev) throw Error(WS_ERR_NOCONV, "eigensolver: " + std::to_string(nconv) + " of " + std::to_string(nev) +
                                                        " eigenpairs converged in " + std::to_string(restarts) + " restarts");
    }
};

// argument checks and defaults shared by the single and the batched entry point; returns false when the problem goes
// to the block-2 iteration instead
void eig_configure(EigRun& R, const ws_pattern* p, ws_solver* s, const double* B_vals, double sigma, int kind, int nev, int ncv,
                   double tol, int maxit, uint64_t seed, const int32_t* edges, const double* xy, int64_t Ntt, double* w_out,
                   double* V_out, int device, cudaStream_t st) {
    WS_REQUIRE(p && s && B_vals && w_out && V_out, "null pointer");
    WS_REQUIRE(kind == 0 || kind == 1, "kind must be 0 (B-orthogonal, symmetric) or 1 (Euclidean, non-symmetric)");
    const int64_t N = p->N;
    WS_REQUIRE(nev >= 1 && nev < N - 1, "1 <= nev < N-1");
    // scipy's default is ncv = max(2k+1, 20) (arpack.py); with the convergence test at tol = eps that size sits on a
    // cliff for these problems (86 or 97 operator applications for k = 10 depending on rounding noise), three more
    // vectors make it 80 for every mesh tried (tools/ncv_sweep.py).
    // Wide bases: the reference's own calls go up to Nmax = 100 (lantern-demo.ipynb cell 2; cell 4 sweeps with
    // Nmax = 80), i.e. ncv = 204; beyond MAX_NCV the default is clamped (fewer spare directions, more restarts).
    if (ncv <= 0) ncv = std::min(std::max(2 * nev + 4, 20), std::max(MAX_NCV, nev + 2));
    ncv = (int)std::min<int64_t>(ncv, N);
    WS_REQUIRE(ncv > nev + 1 && ncv <= MAX_NCV, "nev+1 < ncv <= 256 (Nmax <= 254)");
    WS_REQUIRE(kind == 0 || (edges && xy), "the vector eigensolver needs the edge table and coordinates");
    if (tol <= 0) tol = 2.220446049250313e-16;          // ARPACK: tol=0 -> machine epsilon
    if (maxit <= 0) maxit = 300;
    {
        static PerDeviceOnce configure;
        configure([] { WS_CUDA(cudaFuncSetAttribute(k_update_dots, cudaFuncAttributeMaxDynamicSharedMemorySize, (MAXV + 1) * UD_TILE * (int)sizeof(double))); });
    }
    R.p = p; R.s = s; R.B_vals = B_vals; R.sigma = sigma; R.kind = kind; R.nev = nev; R.m = ncv; R.tol = tol; R.maxit = maxit;
    R.seed = seed; R.edges = edges; R.xy = xy; R.Ntt = Ntt; R.w_out = w_out; R.V_out = V_out; R.device = device; R.st = st;
}

}  // namespace

struct ws_eig_workspace {
    EigRun run;
};
static_assert(sizeof(ws_eig_job) == 400, "ws_eig_job layout is mirrored by wavesolve_b200/_lib.py:EigJob");

int ws_eig_solve(const ws_pattern* p, ws_solver* s, const double* B_vals, double sigma, int kind, int nev, int ncv,
                 double tol, int maxit, uint64_t seed, const int32_t* edges, const double* xy, int64_t Ntt,
                 double* w_out, double* V_out, int64_t* h_info, double* h_resid, int device, void* stream) {
    WS_API_BEGIN
    WS_NVTX("ws K5-K10 Krylov iterate");
    WS_REQUIRE(kind == 0 || kind == 1 || kind == 2,
               "kind must be 0 (B-orthogonal, symmetric), 1 (Euclidean, non-symmetric) or 2 (as 1, block size 2)");
    DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    {
        static const int block = getenv("WS_EIG_BLOCK") ? atoi(getenv("WS_EIG_BLOCK")) : 1;
        if (kind == 2 || (kind == 1 && block == 2)) {
            WS_REQUIRE(p && s && B_vals && w_out && V_out && edges && xy, "null pointer");
            int nc = ncv <= 0 ? std::max(2 * nev + 4, 20) : ncv;
            nc = (int)std::min<int64_t>(nc, p->N);
            if (p->N > 4 * nc) {
                int mb = nc + (nc & 1);
                if (mb + 2 >= MAXV) mb -= 2;
                WS_REQUIRE(mb + 2 < MAXV && mb > nev + 1, "the block-2 iteration (kind 2) handles bases below 62 vectors (Nmax <= 28)");
                if (tol <= 0) tol = 2.220446049250313e-16;
                if (maxit <= 0) maxit = 300;
                eig_solve_block2(p, s, B_vals, sigma, nev, mb, tol, maxit, seed, edges, xy, Ntt, w_out, V_out, h_info, h_resid,
                                 device, st);
                return WS_OK;
            }
            kind = 1;
        }
    }
    EigRun R;
    eig_configure(R, p, s, B_vals, sigma, kind, nev, ncv, tol, maxit, seed, edges, xy, Ntt, w_out, V_out, device, st);
    R.init();
    do {
        R.enqueue_cycle();
    } while (!R.finish_cycle());
    R.enqueue_results();
    R.finish(h_info, h_resid);
    WS_API_END
}

// Several eigenproblems in lockstep, one per stream, driven by this one host thread (see EigRun).  A job that fails
// gets its own rc / message; the others run to the end.  Returns WS_OK unless the arguments themselves are bad.
int ws_eig_solve_batch(ws_eig_job* jobs, int njobs, int device) {
    WS_API_BEGIN
    WS_NVTX("ws K5-K10 Krylov iterate (batch)");
    WS_REQUIRE(jobs && njobs >= 1, "bad arguments");
    DeviceGuard g(device);
    std::vector<std::unique_ptr<EigRun>> owned((size_t)njobs);
    std::vector<EigRun*> runs((size_t)njobs, nullptr);
    std::vector<int> state((size_t)njobs, 0);      // 0 iterating, 1 results enqueued, 2 finished / failed
    auto guarded = [&](int i, auto&& fn) {
        try {
            fn();
            return true;
        } catch (const Error& e) {
            jobs[i].rc = e.code;
            snprintf(jobs[i].error, sizeof jobs[i].error, "%s", e.what());
        } catch (const std::exception& e) {
            jobs[i].rc = WS_ERR_CUDA;
            snprintf(jobs[i].error, sizeof jobs[i].error, "%s", e.what());
        }
        state[i] = 2;
        return false;
    };
    for (int i = 0; i < njobs; ++i) {
        ws_eig_job& J = jobs[i];
        J.rc = WS_OK;
        J.error[0] = 0;
        if (J.workspace) {
            runs[i] = &J.workspace->run;                 // persistent: basis, pinned buffers and step graphs are reused
        } else {
            owned[i].reset(new EigRun());
            runs[i] = owned[i].get();
        }
        guarded(i, [&] {
            eig_configure(*runs[i], J.pattern, J.solver, J.B_vals, J.sigma, J.kind, J.nev, J.ncv, J.tol, J.maxit, J.seed, J.edges,
                          J.xy, J.Ntt, J.w, J.V, device, (cudaStream_t)J.stream);
            runs[i]->init();
        });
    }
    // Event-driven, not lockstep: whichever job's restart cycle has finished is analysed and gets its next cycle
    // enqueued at once, so the jobs drift out of phase and the device always has the other jobs' cycles queued while
    // the host works on one (a barrier per cycle left the GPU idle for the host analysis of every job in turn).
    const bool btrace = getenv("WS_BATCH_TRACE") != nullptr;
    const double tb0 = EigRun::now();
    int ncycles = 0, active = 0;
    for (int i = 0; i < njobs; ++i)
        if (state[i] == 0 && guarded(i, [&] { runs[i]->enqueue_cycle(); })) ++active;
    static const bool lockstep = getenv("WS_BATCH_LOCKSTEP") && atoi(getenv("WS_BATCH_LOCKSTEP"));
    while (active > 0 && lockstep) {
        // (A/B switch: barrier per cycle -- wait for every job, then analyse and re-enqueue every job)
        for (int i = 0; i < njobs; ++i)
            if (state[i] == 0) {
                ++ncycles;
                const bool ok = guarded(i, [&] {
                    if (runs[i]->finish_cycle()) {
                        runs[i]->enqueue_results();
                        state[i] = 1;
                    }
                });
                if (!ok || state[i] != 0) --active;
            }
        for (int i = 0; i < njobs; ++i)
            if (state[i] == 0 && !guarded(i, [&] { runs[i]->enqueue_cycle(); })) --active;
    }
    while (active > 0) {
        bool progressed = false;
        for (int i = 0; i < njobs; ++i) {
            if (state[i] != 0) continue;
            const cudaError_t q = cudaEventQuery(runs[i]->ev);
            if (q == cudaErrorNotReady) continue;
            progressed = true;
            ++ncycles;
            const bool ok = guarded(i, [&] {
                if (q != cudaSuccess) throw Error(WS_ERR_CUDA, std::string("cudaEventQuery: ") + cudaGetErrorString(q));
                if (runs[i]->finish_cycle()) {
                    runs[i]->enqueue_results();
                    state[i] = 1;
                } else {
                    runs[i]->enqueue_cycle();
                }
            });
            if (!ok || state[i] != 0) --active;
        }
        if (!progressed) std::this_thread::yield();
    }
    if (btrace) {
        double tc = 0;
        int nc = 0;
        for (int i = 0; i < njobs; ++i) { tc += runs[i]->t_capture; nc += runs[i]->n_capture; }
        fprintf(stderr, "[ws_eig_solve_batch] %d jobs, %d cycles, %.1f ms; graph captures so far in these work spaces: %d, %.1f ms\n", njobs,
                ncycles, (EigRun::now() - tb0) * 1e3, nc, tc * 1e3);
    }
    for (int i = 0; i < njobs; ++i)
        if (state[i] == 1) {
            if (guarded(i, [&] { runs[i]->finish(jobs[i].info, jobs[i].resid); })) state[i] = 2;
            else if (jobs[i].rc == WS_ERR_NOCONV) {   // results are written: report what the iteration had
                jobs[i].info[0] = runs[i]->nconv; jobs[i].info[1] = runs[i]->nops; jobs[i].info[2] = runs[i]->restarts;
            }
        }
    owned.clear();
    WS_API_END
}

int ws_eig_workspace_create(ws_eig_workspace** out) {
    WS_API_BEGIN
    WS_REQUIRE(out, "out == NULL");
    *out = new ws_eig_workspace();
    WS_API_END
}

int ws_eig_workspace_destroy(ws_eig_workspace* w) {
    WS_API_BEGIN
    delete w;
    WS_API_END
}

}  // extern "C"
