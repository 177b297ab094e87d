// Host-side construction of the filter instances that the reference obtains from RadiationDetectorDSP.jl's
// fltinstance(...) (call sites: /root/reference/src/dsp_icpc.jl:157-181, src/dsp_routines.jl:12,56).
// Pure CPU code, no CUDA; exported through the C ABI (include/lgdsp_b200.h).
//
// Numerics differ on purpose from the oracle's independent builders (oracle/lgdsp_oracle.c): here the
// least-squares problems are solved with orthogonal (discrete Legendre / Gram) polynomials in long double,
// the oracle uses normal equations + Gauss-Jordan.  tests/test_builders.py checks they agree to 1e-12.
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>
#include "../../include/lgdsp_b200.h"

namespace {

// Gram polynomial basis on the points x_i = i - c, i = 0..n-1, built by Gram-Schmidt on monomials.
// Returns Q[j][i] (orthonormal, j = 0..degree) and the monomial expansion R such that
// q_j(x) = sum_{p<=j} R[j][p] * x^p.
struct GramBasis {
    int n, m;
    std::vector<long double> Q;  // m x n
    std::vector<long double> R;  // m x m (lower triangular), monomial coefficients
};

GramBasis gram_basis(int n, int degree, long double c)
{
    GramBasis B;
    B.n = n;
    B.m = degree + 1;
    B.Q.assign((size_t)B.m * n, 0.0L);
    B.R.assign((size_t)B.m * B.m, 0.0L);
    std::vector<long double> x(n);
    for (int i = 0; i < n; ++i) x[i] = (long double)i - c;
    for (int j = 0; j < B.m; ++j) {
        // start with the monomial x^j
        std::vector<long double> v(n), coef(B.m, 0.0L);
        for (int i = 0; i < n; ++i) v[i] = powl(x[i], j);
        coef[j] = 1.0L;
        // two rounds of Gram-Schmidt for orthogonality
        for (int round = 0; round < 2; ++round)
            for (int p = 0; p < j; ++p) {
                long double dot = 0;
                for (int i = 0; i < n; ++i) dot += v[i] * B.Q[(size_t)p * n + i];
                for (int i = 0; i < n; ++i) v[i] -= dot * B.Q[(size_t)p * n + i];
                for (int s = 0; s <= p; ++s) coef[s] -= dot * B.R[(size_t)p * B.m + s];
            }
        long double nrm = 0;
        for (int i = 0; i < n; ++i) nrm += v[i] * v[i];
        nrm = sqrtl(nrm);
        for (int i = 0; i < n; ++i) B.Q[(size_t)j * n + i] = v[i] / nrm;
        for (int s = 0; s <= j; ++s) B.R[(size_t)j * B.m + s] = coef[s] / nrm;
    }
    return B;
}

// monomial-coefficient fit matrix about centre c: coef_p = sum_i A[i][p] y_i, polynomial in (x - c)
void fit_matrix_centered(int n, int degree, long double c, std::vector<long double>& A)
{
    GramBasis B = gram_basis(n, degree, c);
    int m = B.m;
    A.assign((size_t)n * m, 0.0L);
    // y ~ sum_j (q_j . y) q_j(x) = sum_j (q_j . y) sum_p R[j][p] x^p
    for (int i = 0; i < n; ++i)
        for (int p = 0; p < m; ++p) {
            long double s = 0;
            for (int j = p; j < m; ++j) s += B.Q[(size_t)j * n + i] * B.R[(size_t)j * m + p];
            A[(size_t)i * m + p] = s;
        }
}

void cuspzac_shape(double sigma, int flat, int L, bool zac, std::vector<long double>& c)
{
    const int lt = (L - flat) / 2;
    const long double s = sigma, h = lt / 2.0L;
    const long double norm = sinhl(lt / s);
    c.assign(L, 0.0L);
    std::vector<long double> par(L, 0.0L);
    long double acusp = 0, apar = 0;
    for (int k = 0; k < L; ++k) {
        if (k < lt) {
            c[k] = sinhl(k / s) / norm;
            par[k] = (k - h) * (k - h) - h * h;
        } else if (k <= lt + flat) {
            c[k] = 1.0L;
        } else {
            c[k] = sinhl((L - k) / s) / norm;
            par[k] = (L - k - h) * (L - k - h) - h * h;
        }
        acusp += c[k];
        apar += par[k];
    }
    if (zac && apar != 0.0L)
        for (int k = 0; k < L; ++k) c[k] -= par[k] / apar * acusp;
}

int cuspzac_coeffs(double sigma, int flat, double tau, int L, double beta, bool zac, double* out)
{
    if (!out || L < 4 || L > LGDSP_MAX_FIR || flat < 0 || flat >= L - 2 || !(sigma > 0) || !(tau > 0))
        return LGDSP_ERR_INVALID_ARG;
    std::vector<long double> c;
    cuspzac_shape(sigma, flat, L, zac, c);
    const long double r = expl(-1.0L / (long double)tau);
    const long double g = (long double)beta / L;
    for (int k = 0; k < L; ++k) {
        long double v = c[k] - (k > 0 ? r * c[k - 1] : 0.0L);
        out[k] = (double)(v * g);
    }
    return LGDSP_OK;
}

}  // namespace

extern "C" {

// [RDDSP _lsq_fit_matrix(0:n-1, degree)]; usage /root/reference/src/multi_intersect.jl:80,115-119
int lgdsp_lsq_fit_matrix(int32_t n, int32_t degree, double* A)
{
    if (!A || degree < 0 || degree > 7 || n <= degree || n > 4096) return LGDSP_ERR_INVALID_ARG;
    // fit about the window centre for conditioning, then re-expand the polynomial about x = 0
    const long double c = (n - 1) / 2.0L;
    std::vector<long double> Ac;
    fit_matrix_centered(n, degree, c, Ac);
    const int m = degree + 1;
    // (x - c)^p = sum_q binom(p,q) x^q (-c)^(p-q)
    std::vector<long double> binom((size_t)m * m, 0.0L);
    for (int p = 0; p < m; ++p) {
        binom[(size_t)p * m] = 1.0L;
        for (int q = 1; q <= p; ++q)
            binom[(size_t)p * m + q] = binom[(size_t)(p - 1) * m + q - 1] + (q <= p - 1 ? binom[(size_t)(p - 1) * m + q] : 0.0L);
    }
    for (int i = 0; i < n; ++i)
        for (int q = 0; q < m; ++q) {
            long double s = 0;
            for (int p = q; p < m; ++p) s += Ac[(size_t)i * m + p] * binom[(size_t)p * m + q] * powl(-c, p - q);
            A[(size_t)i * m + q] = (double)s;
        }
    return LGDSP_OK;
}

// [RDDSP SavitzkyGolayFilter(length, degree, derivative)] coefficients, s[j] = sum_k h[k] y[j+k]
int lgdsp_sg_coeffs(int32_t n_taps, int32_t degree, int32_t derivative, double* h)
{
    if (!h || n_taps < 1 || (n_taps & 1) == 0 || n_taps > 4095 || degree < 0 || degree > 7 || derivative < 0 ||
        derivative > degree)
        return LGDSP_ERR_INVALID_ARG;
    if (n_taps <= degree) {
        // fewer points than polynomial coefficients (dsp_icpc_compressed's in-trace filter with the example config:
        // 3 taps, degree 3, /root/reference/src/dsp_icpc.jl:439): minimum-norm solution of the underdetermined fit,
        // c = V^T (V V^T)^-1 y with V[k][p] = x_k^p, so h = derivative! * (V V^T)^-1 V[:, derivative]  (policy, unpinned)
        const int n = n_taps, m = degree + 1;
        long double G[8][9];
        for (int i = 0; i < n; ++i) {
            const long double xi = (long double)(i - n / 2);
            for (int k = 0; k < n; ++k) {
                const long double xk = (long double)(k - n / 2);
                long double g = 0;
                for (int p = 0; p < m; ++p) g += powl(xi, p) * powl(xk, p);
                G[i][k] = g;
            }
            G[i][n] = powl(xi, derivative);
        }
        for (int c = 0; c < n; ++c) {   // Gauss-Jordan with partial pivoting
            int piv = c;
            for (int r = c + 1; r < n; ++r) if (fabsl(G[r][c]) > fabsl(G[piv][c])) piv = r;
            if (G[piv][c] == 0.0L) return LGDSP_ERR_INVALID_ARG;
            if (piv != c) for (int k = 0; k <= n; ++k) std::swap(G[piv][k], G[c][k]);
            for (int r = 0; r < n; ++r) {
                if (r == c) continue;
                const long double f = G[r][c] / G[c][c];
                for (int k = c; k <= n; ++k) G[r][k] -= f * G[c][k];
            }
        }
        long double fact = 1.0L;
        for (int k = 2; k <= derivative; ++k) fact *= k;
        for (int i = 0; i < n; ++i) h[i] = (double)(fact * G[i][n] / G[i][i]);
        return LGDSP_OK;
    }
    std::vector<long double> Ac;
    fit_matrix_centered(n_taps, degree, (long double)(n_taps / 2), Ac);
    long double fact = 1.0L;
    for (int k = 2; k <= derivative; ++k) fact *= k;
    const int m = degree + 1;
    for (int i = 0; i < n_taps; ++i) h[i] = (double)(fact * Ac[(size_t)i * m + derivative]);
    return LGDSP_OK;
}

int lgdsp_cusp_coeffs(double sigma, int32_t flat, double tau, int32_t n_taps, double beta, double* c)
{
    return cuspzac_coeffs(sigma, flat, tau, n_taps, beta, false, c);
}

int lgdsp_zac_coeffs(double sigma, int32_t flat, double tau, int32_t n_taps, double beta, double* c)
{
    return cuspzac_coeffs(sigma, flat, tau, n_taps, beta, true, c);
}

}  // extern "C"
