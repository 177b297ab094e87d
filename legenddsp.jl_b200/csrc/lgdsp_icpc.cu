// Fused dsp_icpc kernel for sm_100a: one CTA (256 threads) per waveform, persistent over the event slice.
//
// Data flow per waveform (reference steps in brackets, /root/reference/src/dsp_icpc.jl):
//   TMA bulk copy (cp.async.bulk, 16 KB UInt16) HBM -> SMEM, prefetched one event ahead
//   pass 1  raw samples: saturation [:93-95], baseline regression sums [:102], min/max [:111-112],
//           integer prefix sums P = cumsum(x), PP = cumsum(P)            (exact integer arithmetic)
//   pass 2  pole-zero in closed form y = w + km1*cumsum(w), w = x - blmean [:105,:119-120]; writes
//           TT[i+1] = cumsum(y)[i] (float64, SMEM); tail log-regression [:115]; PZ tail stats [:123];
//           threshold masks for t10..t99 [:132-136]
//   pass 3  every trapezoid [:126,:147-164,:202-207] as 4 look-ups in TT per output; Savitzky-Golay and
//           derivative currents [:181-186] as sliding-window FIRs on TT
//   pass 4  masks for t50_current and the in-trace pile-up search [:189-195]; crossing resolution with
//           bit-parallel run detection (Intersect state machine)
//   cz      CUSP/ZAC [:167-178] through their analytic structure (sliding exponential/polynomial windows)
//   pass 5  interpolated pick-offs (PolynomialDNI), qdrift/lq [:141-144], output row (49 doubles)
//
// Work mapping: thread t owns the CH = 33 consecutive samples/outputs [33t, 33t+33).  33 is odd, so with the plain
// linear layout of TT the 32 lanes of a warp hit 32 different banks for EVERY window offset, and all SMEM
// addresses inside the unrolled loops are "per-stream base register + immediate": no padding, no index math.
// The waveform is read from HBM exactly once (16 KB) and 392 B are written.
#include <cuda_runtime.h>
#include <math_constants.h>
#include "lgdsp_device.cuh"
#include "lgdsp_kernels.h"

namespace lgdsp {

constexpr int NT = 256;          // threads per CTA
constexpr int NWARP = NT / 32;
constexpr int CH = CZ_CH;        // samples per thread (33)
constexpr int MAXN = LGDSP_MAX_SAMPLES;
constexpr int NWORDS = MAXN / 32;  // mask words
static_assert(NT * CH >= MAXN + 1, "chunks must cover the waveform");
static_assert(NWORDS == NT, "one mask word per thread");

enum { M_T0 = 0, M_T0INV, M_T10, M_T50, M_T80, M_T90, M_T99, M_CUR, M_PILE, NMASK };

// ---- shared memory carve-up (bytes) ----
constexpr int SM_XS = 0;                                   // uint16 xs[8192]  (aliased by CUSP/ZAC tables 0..7)
constexpr int TT_LEN = MAXN + 8;
constexpr int SM_TT = SM_XS + MAXN * 2;                    // double TT[8193+]
constexpr int SM_MASK = SM_TT + TT_LEN * 8;                // uint32 masks[NMASK][NWORDS]
constexpr int RED_W = 24;
constexpr int SM_RED = SM_MASK + NMASK * NWORDS * 4;       // double red[NWARP][24]
constexpr int SM_STASH = SM_RED + NWARP * RED_W * 8;       // double stash[3][LGDSP_MAX_DNI]
constexpr int SM_TAB = SM_STASH + 3 * LGDSP_MAX_DNI * 8;   // double tabB[8][256]: CUSP/ZAC prefix tables 8..15
constexpr int SM_ROW = SM_TAB + 8 * NT * 8;                // double row[64]
constexpr int SM_SCR = SM_ROW + 64 * 8;                    // double scr[32]: scalars passed between warps
constexpr int SM_IBUF = SM_SCR + 32 * 8;                   // int ibuf[64]
constexpr int SM_BAR = SM_IBUF + 64 * 4;                   // uint64 mbarrier
constexpr int SM_TOTAL = SM_BAR + 16;

int icpc_smem_bytes() { return SM_TOTAL; }
int icpc_threads() { return NT; }

enum { IB_POS0 = 0 /* NMASK positions */, IB_MULT = 16, IB_SCAN = 32 /* 8 warp totals */ };
enum { SC_TX = 0 /* 5 */, SC_T0 = 5, SC_T0INV = 6, SC_DSCAN = 8 /* 8 */, SC_PP0 = 16 };

// ---------------------------------------------------------------------------------------------------
// block-wide reductions; every thread gets the result
// ---------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red, int tid)
{
    static_assert(NV <= RED_W, "reduction scratch too small");
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[wid * RED_W + i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) s += red[w * RED_W + i];
        v[i] = s;
    }
    __syncthreads();
}

template <int NV>
__device__ __forceinline__ void block_max(double (&v)[NV], double* red, int tid)
{
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_max(v[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[wid * RED_W + i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = red[i];
#pragma unroll
        for (int w = 1; w < NWARP; ++w) s = fmax(s, red[w * RED_W + i]);
        v[i] = s;
    }
    __syncthreads();
}

// block-wide (max value, FIRST index) for NV candidates
template <int NV>
__device__ __forceinline__ void block_argmax(double (&v)[NV], int (&ix)[NV], double* red, int tid)
{
    static_assert(2 * NV <= RED_W, "reduction scratch too small");
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) warp_argmax(v[i], ix[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            red[wid * RED_W + 2 * i] = v[i];
            red[wid * RED_W + 2 * i + 1] = (double)ix[i];
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double bv = red[2 * i];
        int bi = (int)red[2 * i + 1];
#pragma unroll
        for (int w = 1; w < NWARP; ++w) {
            double ov = red[w * RED_W + 2 * i];
            int oi = (int)red[w * RED_W + 2 * i + 1];
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        v[i] = bv;
        ix[i] = bi;
    }
    __syncthreads();
}

// run-length monoid for the saturation counters (src/saturation.jl:28-65)
struct Run {
    int pre, suf, best, len;
};
__device__ __forceinline__ Run run_merge(const Run& A, const Run& B)
{
    Run r;
    r.pre = (A.pre == A.len) ? A.len + B.pre : A.pre;
    r.suf = (B.suf == B.len) ? B.len + A.suf : B.suf;
    r.best = max(max(A.best, B.best), A.suf + B.pre);
    r.len = A.len + B.len;
    return r;
}
__device__ __forceinline__ Run run_shfl_down(const Run& a, int o)
{
    Run r;
    r.pre = __shfl_down_sync(FULL, a.pre, o);
    r.suf = __shfl_down_sync(FULL, a.suf, o);
    r.best = __shfl_down_sync(FULL, a.best, o);
    r.len = __shfl_down_sync(FULL, a.len, o);
    return r;
}

// regression statistics from accumulated sums, operation order of the reference (src/tailstats.jl:54-70 and
// its RDDSP original signalstats); never contracted
struct Stats {
    double mean, sigma, slope, offset;
};
__device__ __noinline__ Stats stats_finalize(int n, double sX, double sXX, double sY, double sYY, double sXY)
{
    const double inv_n = div_rn(1.0, (double)n);
    const double mean_X = mul_rn(sX, inv_n);
    const double mean_Y = mul_rn(sY, inv_n);
    const double var_X = sub_rn(mul_rn(sXX, inv_n), mul_rn(mean_X, mean_X));
    double var_Y = sub_rn(mul_rn(sYY, inv_n), mul_rn(mean_Y, mean_Y));
    const double cov = sub_rn(mul_rn(sXY, inv_n), mul_rn(mean_X, mean_Y));
    Stats s;
    s.slope = div_rn(cov, var_X);
    s.offset = sub_rn(mean_Y, mul_rn(s.slope, mean_X));
    if (var_Y < 0) var_Y = 0;
    s.mean = mean_Y;
    s.sigma = sqrt(var_Y);
    return s;
}

// sum_{i=a}^{b} X_i and X_i^2 with X_i = t0 + i*dt
__device__ __noinline__ void xsums(int a, int b, double t0, double dt, double& sX, double& sXX)
{
    const double cnt = (double)(b - a + 1);
    const double si = 0.5 * (double)(a + b) * cnt;
    auto s2 = [](double k) { return k * (k + 1.0) * (2.0 * k + 1.0) / 6.0; };
    const double sii = s2((double)b) - s2((double)a - 1.0);
    sX = cnt * t0 + dt * si;
    sXX = cnt * t0 * t0 + 2.0 * t0 * dt * si + dt * dt * sii;
}

// extrema3points  src/interpolation.jl:8-10
__device__ __forceinline__ double extrema3(double y1, double y2, double y3)
{
    const double a = y3 - 4.0 * y2 + 3.0 * y1;
    return y1 - a * a / (8.0 * (y3 - 2.0 * y2 + y1));
}

// ---- single-sample evaluations on the prefix sums TT (TT[i] = sum_{k<i} y[k]); used by the scalar tail ----
__device__ __forceinline__ double trap_at(const double* TT, const TrapDev& t, int j)
{
    const double s1 = TT[j + t.a] - TT[j];
    const double s2 = TT[j + t.L] - TT[j + t.a + t.g];
    return s2 * t.inv2 - s1 * t.inv1;
}
__device__ __forceinline__ double y_at(const double* TT, int i) { return TT[i + 1] - TT[i]; }
__device__ __noinline__ double sg_at(const double* TT, const SgDev& s, int j)
{
    double a0 = 0, a1 = 0;   // same association as sg_chunk_t
#pragma unroll 1
    for (int k = 0; k <= s.n_taps; k += 2) {
        a0 = fma(s.gg[k], TT[j + k], a0);
        if (k + 1 <= s.n_taps) a1 = fma(s.gg[k + 1], TT[j + k + 1], a1);
    }
    return a0 + a1;
}
// DerivativeFilter sample i  (src/derivative.jl:47-55)
__device__ __forceinline__ double deriv_at(const double* TT, int i)
{
    const int ii = i < 1 ? 1 : i;
    return (TT[ii + 1] - TT[ii]) - (TT[ii] - TT[ii - 1]);
}
// direct FIR output j of a CUSP/ZAC filter with differenced taps g[0..L] on TT: out[j] = sum_k g[k] TT[j+L-k]
__device__ __forceinline__ double fir_at(const double* TT, const double* __restrict__ g, int L, int j)
{
    double acc = 0;
    for (int k = 0; k <= L; ++k) acc = fma(__ldg(g + k), TT[j + L - k], acc);
    return acc;
}

// thread-local 33-bit mask -> block mask (bit position 33*tid + k)
__device__ __forceinline__ void mask_commit(uint32_t* M, int tid, unsigned long long bits)
{
    if (bits == 0ull) return;
    const int p0 = tid * CH, w0 = p0 >> 5, s0 = p0 & 31;
    const unsigned long long sh = bits << s0;   // 33 bits shifted by <= 31: fits in 64
    const uint32_t lo = (uint32_t)sh, mid = (uint32_t)(sh >> 32);
    if (lo) atomicOr(&M[w0], lo);
    if (mid && w0 + 1 < NWORDS) atomicOr(&M[w0 + 1], mid);
}
// same for the time-reversed trace of length nlen: forward bit i <-> reversed bit nlen-1-i
__device__ __forceinline__ void mask_commit_reversed(uint32_t* M, int tid, unsigned long long bits, int nlen)
{
    if (bits == 0ull) return;
    // reversed positions of the chunk: base = nlen-1-(33*tid+32) holds forward bit 32, base+32 holds forward bit 0
    unsigned long long rev = __brevll(bits) >> (64 - CH);   // rev bit (32-k) = bits bit k
    int base = nlen - 1 - (tid * CH + CH - 1);
    if (base < 0) { rev >>= (-base); base = 0; }
    const int w0 = base >> 5, s0 = base & 31;
    const unsigned long long sh = rev << s0;
    const uint32_t lo = (uint32_t)sh, mid = (uint32_t)(sh >> 32);
    if (lo) atomicOr(&M[w0], lo);
    if (mid && w0 + 1 < NWORDS) atomicOr(&M[w0 + 1], mid);
}

// One warp: find runs of >= k consecutive set bits in the NWORDS-word mask M (bits beyond the trace are zero)
// that do not start at bit 0 -- the Intersect state machine: `pos` = start of the first such
// run (-1 if none), `mult` = number of such runs.  M is destroyed.
__device__ __noinline__ void resolve_runs(uint32_t* M, int k, int lane, int& pos, int& mult)
{
    constexpr int Q = NWORDS / 32;
    uint32_t m[Q], r[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) { m[q] = M[lane * Q + q]; r[q] = m[q]; }
    int len = 1;
    while (len < k) {
        const int step = min(len, k - len);
        const int ws = step >> 5, bs = step & 31;
        __syncwarp();
        uint32_t nr[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int w = lane * Q + q + ws;
            const uint32_t lo = w < NWORDS ? M[w] : 0u;
            const uint32_t hi = (w + 1) < NWORDS ? M[w + 1] : 0u;
            const uint32_t sh = bs ? ((lo >> bs) | (hi << (32 - bs))) : lo;
            nr[q] = r[q] & sh;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < Q; ++q) { r[q] = nr[q]; M[lane * Q + q] = nr[q]; }
        len += step;
    }
    uint32_t prev_top = __shfl_up_sync(FULL, m[Q - 1], 1);
    if (lane == 0) prev_top = 0;
    int p = 0x7fffffff, cnt = 0;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const uint32_t carry = ((q == 0) ? prev_top : m[q - 1]) >> 31;
        uint32_t c = r[q] & ~((m[q] << 1) | carry);
        if (lane == 0 && q == 0) c &= ~1u;  // a run that starts at the first sample never fires
        cnt += __popc(c);
        if (c && p == 0x7fffffff) p = (lane * Q + q) * 32 + (__ffs(c) - 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        p = min(p, __shfl_xor_sync(FULL, p, o));
        cnt += __shfl_xor_sync(FULL, cnt, o);
    }
    pos = (p == 0x7fffffff) ? -1 : p;
    mult = cnt;
}

// linear interpolation of Intersect: x = (thr - y_l)*(x_r - x_l)/(y_r - y_l) + x_l
__device__ __noinline__ double cross_x(double thr, double yl, double yr, double tl, double dt)
{
    return (thr - yl) * dt / (yr - yl) + tl;
}

// window placement policy (include/lgdsp_b200.h, lgdsp_dni)
__device__ __noinline__ void dni_window(int n_w, int n_trace, double p, double& pc, int& from)
{
    if (!(p >= 0)) p = 0;
    if (p > n_trace - 1) p = n_trace - 1;
    long long f = (long long)rint(p) - n_w / 2;
    if (f < 0) f = 0;
    if (f > n_trace - n_w) f = n_trace - n_w;
    pc = p;
    from = (int)f;
}
// PolynomialDNI estimate, one warp: window values win[0..n_w) (SMEM/any), fit matrix A (global); result in every lane
__device__ __noinline__ double dni_eval_warp(const double* __restrict__ A, int n_w, int m, const double* win, double u,
                                                int lane)
{
    double c[LGDSP_MAX_DNI_DEG + 1] = {0, 0, 0, 0};
    for (int i = lane; i < n_w; i += 32) {
        const double v = win[i];
#pragma unroll
        for (int j = 0; j <= LGDSP_MAX_DNI_DEG; ++j)
            if (j < m) c[j] = fma(__ldg(A + i * m + j), v, c[j]);
    }
#pragma unroll
    for (int j = 0; j <= LGDSP_MAX_DNI_DEG; ++j) c[j] = warp_sum(c[j]);
    double r = c[m - 1];
    for (int j = m - 2; j >= 0; --j) r = r * u + c[j];
    return r;
}

// ---------------------------------------------------------------------------------------------------
// chunk passes over the prefix sums
// ---------------------------------------------------------------------------------------------------
// one trapezoid over the thread's chunk of outputs j in [j0, j0+CH): max, optional min (as max of -o), optional
// first argmax, optional threshold masks for o >= thr and -o >= thr
template <bool WANT_NEG, bool WANT_ARG, bool WANT_MASK>
__device__ __forceinline__ void trap_chunk(const double* TT, const TrapDev& t, int j0, double thr, double& vmax, double& vneg,
                                           int& arg, unsigned long long& bpos, unsigned long long& bneg)
{
    const int cnt = min(CH, t.nout - j0);
    if (cnt <= 0) return;
    const double* p0 = TT + j0;
    const double* p1 = p0 + t.a;
    const double* p2 = p1 + t.g;
    const double* p3 = p0 + t.L;
    const double inv1 = t.inv1, inv2 = t.inv2;
    auto body = [&](int k) {
        const double o = (p3[k] - p2[k]) * inv2 - (p1[k] - p0[k]) * inv1;
        if (WANT_ARG) {
            if (o > vmax) { vmax = o; arg = j0 + k; }
        } else {
            vmax = fmax(vmax, o);
        }
        if (WANT_NEG) vneg = fmax(vneg, -o);
        if (WANT_MASK) {
            bpos |= (o >= thr) ? (1ull << k) : 0ull;
            bneg |= (-o >= thr) ? (1ull << k) : 0ull;
        }
    };
    // modest unrolling only: the kernel must stay small enough for the instruction cache
    int k = 0;
#pragma unroll 1
    for (; k + 3 <= cnt; k += 3) {
        body(k);
        body(k + 1);
        body(k + 2);
    }
#pragma unroll 1
    for (; k < cnt; ++k) body(k);
}

// sliding-window FIR on TT (the SG kernels folded onto the prefix sums): s[j] = sum_{q<NW} gg[q]*TT[j+q].
// f(k, s) is called for every output j0+k of the chunk, k < cnt.
template <int NW, typename F>
__device__ __forceinline__ void sg_chunk_t(const double* TT, const SgDev& S, int j0, int cnt, F&& f)
{
    double g[NW], w[NW];
#pragma unroll
    for (int q = 0; q < NW; ++q) g[q] = S.gg[q];
    const double* p = TT + j0;
#pragma unroll
    for (int q = 0; q < NW - 1; ++q) w[q] = p[q];
    int k = 0;
    // groups of NW outputs: the window rotates through the NW registers with compile-time indices (no moves)
#pragma unroll 1
    for (; k + NW <= cnt; k += NW) {
#pragma unroll
        for (int u = 0; u < NW; ++u) {
            w[(u + NW - 1) % NW] = p[k + u + NW - 1];
            double a0 = 0, a1 = 0;   // two accumulators: halves the dependent FMA chain
#pragma unroll
            for (int q = 0; q < NW; q += 2) {
                a0 = fma(g[q], w[(u + q) % NW], a0);
                if (q + 1 < NW) a1 = fma(g[q + 1], w[(u + q + 1) % NW], a1);
            }
            f(k + u, a0 + a1);
        }
    }
#pragma unroll 1
    for (; k < cnt; ++k) {
        double a0 = 0, a1 = 0;
#pragma unroll
        for (int q = 0; q < NW; q += 2) {
            a0 = fma(g[q], p[k + q], a0);
            if (q + 1 < NW) a1 = fma(g[q + 1], p[k + q + 1], a1);
        }
        f(k, a0 + a1);
    }
}
template <typename F>
__device__ __forceinline__ void sg_chunk(const double* TT, const SgDev& S, int j0, int cnt, F&& f)
{
    if (cnt <= 0) return;
    switch (S.n_taps + 1) {
        case 6: sg_chunk_t<6>(TT, S, j0, cnt, f); break;    // 5 taps
        case 8: sg_chunk_t<8>(TT, S, j0, cnt, f); break;    // 7 taps
        default:
            for (int k = 0; k < cnt; ++k) f(k, sg_at(TT, S, j0 + k));
    }
}

// ==================================================================================================
// CUSP/ZAC through their analytic structure (replaces two 2375-tap FIRs = 27.6 M MAC per waveform by O(n) work).
//
//   coeffs[k] = g*(c[k] - r*c[k-1])  =>  out[j] = g*( sum_k c[k]*d[m-k] + r*c[L-1]*y[m-L] ),  m = j+L-1,
//   d[i] = y[i] - r*y[i-1]  (the "current"),   c = sinh flank | flat top | mirrored sinh flank (+ parabolas for ZAC).
//
// Every piece of c is an exponential or a polynomial in k, so sum_k c[k]*d[m-k] splits into sliding windows
//   E-(m) = sum rho^k d, E+(m) = sum rho^-k d, W0/W1/W2 = sum {1,k,k^2} d      (rho = exp(-1/sigma))
// over the left flank, the flat top and the right flank.  Each window obeys a 1-step linear recurrence in m.
// Thread t owns outputs m in [33t, 33t+33): it gets the window states at m = 33t in closed form from
// block-wide prefix scans of d (decayed prefix P-, anti-causal decayed prefix P+, moments D1 = sum i*d,
// D2 = sum i^2*d; D0 = sum d comes from TT directly), which are only ever needed at 4 positions per chunk
// (fixed in-chunk offsets) -> 16 tables x 256 entries instead of 4 full-resolution arrays; then it steps the
// recurrences 33 times.  The growing exponentials are only propagated over 33 samples, so nothing blows up.
// (validated against the direct FIR in tools/proto_cuspzac.py and tests/test_gpu_*.py)
// ==================================================================================================
struct CzState {
    double EmL, EpL, W0L, W1L, W2L, W0F, V0, V1, V2, EpR, EmR;
    bool active;
};

__device__ __forceinline__ double* cz_tab(double* tabA, double* tabB, int idx)
{
    return idx < 8 ? tabA + idx * NT : tabB + (idx - 8) * NT;
}

// block-wide scans of d over the whole waveform; fills the 16 decimated tables and *pp0 = P+[0]
__device__ __noinline__ void cz_scan(const CzDev& Z, const double* TT, int n, int tid, double* tabA, double* tabB, double* red,
                        double* pp0)
{
    const int lane = tid & 31, wid = tid >> 5;
    const int i0 = tid * CH;
    const double r = Z.r, rho = Z.rho;
    double pm = 0, d1 = 0, d2 = 0, pp = 0;
    double cpm[4] = {0, 0, 0, 0}, cd1[4] = {0, 0, 0, 0}, cd2[4] = {0, 0, 0, 0}, cpp[4] = {0, 0, 0, 0};
    if (i0 < n) {
        const int cnt = min(CH, n - i0);
        // forward: P-, D1, D2; captured at the 4 in-chunk offsets (sorted ascending on the host: segments, no
        // per-sample compare); samples beyond the trace contribute d = 0
        {
            const double* p = TT + i0;
            double tcur = p[0];
            double yprev = (i0 >= 1) ? tcur - p[-1] : 0.0;
            double di = (double)i0;
            int k = 0;
#pragma unroll 1
            for (int sgm = 0; sgm < 4; ++sgm) {
                const int kend = Z.oc_sorted[sgm];
#pragma unroll 1
                for (; k <= kend; ++k) {
                    double d = 0.0;
                    if (k < cnt) {
                        const double tnext = p[k + 1];
                        const double y = tnext - tcur;      // exact difference of neighbouring prefix sums
                        d = fma(-r, yprev, y);
                        yprev = y;
                        tcur = tnext;
                    }
                    pm = fma(rho, pm, d);
                    d1 = fma(di, d, d1);
                    d2 = fma(di * di, d, d2);
                    di += 1.0;
                }
                cpm[sgm] = pm; cd1[sgm] = d1; cd2[sgm] = d2;
            }
#pragma unroll 1
            for (; k < CH; ++k) {
                double d = 0.0;
                if (k < cnt) {
                    const double tnext = p[k + 1];
                    const double y = tnext - tcur;
                    d = fma(-r, yprev, y);
                    yprev = y;
                    tcur = tnext;
                }
                pm = fma(rho, pm, d);
                d1 = fma(di, d, d1);
                d2 = fma(di * di, d, d2);
                di += 1.0;
            }
        }
        // backward: P+, captured at the 4 offsets sorted DESCENDING
        {
            const double* p = TT + i0;
            double tn = p[cnt], tc = p[cnt - 1];
            int k = CH - 1;
            auto step = [&](int kk) {
                double d = 0.0;
                if (kk < cnt) {
                    const double tp = (i0 + kk >= 1) ? p[kk - 1] : tc;   // sample 0: y[-1] = 0
                    d = fma(-r, tc - tp, tn - tc);
                    tn = tc;
                    tc = tp;
                }
                pp = fma(rho, pp, d);
            };
#pragma unroll 1
            for (int sgm = 0; sgm < 4; ++sgm) {
                const int kend = Z.oa_sorted[sgm];
#pragma unroll 1
                for (; k >= kend; --k) step(k);
                cpp[sgm] = pp;
            }
#pragma unroll 1
            for (; k >= 0; --k) step(k);
        }
    }
    // warp-level scans (linear recurrences with constant multiplier rho^CH; plain sums for the moments)
    double vpm = pm, vd1 = d1, vd2 = d2, vpp = pp;
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int o = 1 << s;
        const double upm = __shfl_up_sync(FULL, vpm, o), ud1 = __shfl_up_sync(FULL, vd1, o), ud2 = __shfl_up_sync(FULL, vd2, o);
        const double upp = __shfl_down_sync(FULL, vpp, o);
        if (lane >= o) { vpm = fma(Z.rho_ch_pow[s], upm, vpm); vd1 += ud1; vd2 += ud2; }
        if (lane + o < 32) vpp = fma(Z.rho_ch_pow[s], upp, vpp);
    }
    if (lane == 31) { red[wid] = vpm; red[8 + wid] = vd1; red[16 + wid] = vd2; }
    if (lane == 0) red[24 + wid] = vpp;
    __syncthreads();
    double gpm = 0, gd1 = 0, gd2 = 0, gpp = 0;   // carries of the neighbouring warps
    for (int w = 0; w < wid; ++w) { gpm = fma(Z.rho_warp, gpm, red[w]); gd1 += red[8 + w]; gd2 += red[16 + w]; }
    for (int w = NWARP - 1; w > wid; --w) gpp = fma(Z.rho_warp, gpp, red[24 + w]);
    // global inclusive values of this thread, then the carry-in = inclusive value of the neighbour thread
    const double ipm = fma(Z.rho_lane[lane], gpm, vpm);
    const double ipp = fma(Z.rho_lane[31 - lane], gpp, vpp);
    double c_pm = __shfl_up_sync(FULL, ipm, 1), c_pp = __shfl_down_sync(FULL, ipp, 1);
    if (lane == 0) c_pm = gpm;
    if (lane == 31) c_pp = gpp;
    const double c_d1 = gd1 + (vd1 - d1), c_d2 = gd2 + (vd2 - d2);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        // q = rank in the sorted order; the logical table index comes from the host (tab_c/tab_a)
        cz_tab(tabA, tabB, Z.tab_c[q] * 3 + 0)[tid] = fma(Z.pw_c[q], c_pm, cpm[q]);
        cz_tab(tabA, tabB, Z.tab_c[q] * 3 + 1)[tid] = cd1[q] + c_d1;
        cz_tab(tabA, tabB, Z.tab_c[q] * 3 + 2)[tid] = cd2[q] + c_d2;
        cz_tab(tabA, tabB, 12 + Z.tab_a[q])[tid] = fma(Z.pw_a[q], c_pp, cpp[q]);
    }
    if (tid == 0) *pp0 = ipp;
}

// window states at m = CH*tid in closed form from the tables
__device__ __noinline__ void cz_init(const CzDev& Z, const double* TT, int n, int tid, double* tabA, double* tabB, double pp0,
                        CzState& S)
{
    const int m = tid * CH;
    S.active = (m < n) && (m + CH - 1 >= Z.L - 1);
    if (!S.active) return;
    auto lc = [&](int q, int kind, int pos) -> double {   // causal prefixes (P-, D1, D2): zero before the trace
        return pos < 0 ? 0.0 : cz_tab(tabA, tabB, q * 3 + kind)[pos / CH];
    };
    auto la = [&](int q, int pos) -> double {             // anti-causal prefix P+
        if (pos >= n) return 0.0;
        if (pos < 0) return exp((double)pos * Z.inv_sigma) * pp0;   // rho^(-pos) * P+[0]
        return cz_tab(tabA, tabB, 12 + q)[pos / CH];
    };
    auto d0 = [&](int j) -> double {                      // D0[j] = sum_{i<=j} d[i] = TT[j+1] - r*TT[j]
        return j < 0 ? 0.0 : fma(-Z.r, TT[j], TT[j + 1]);
    };
    const int lt = Z.lt, L = Z.L, Rn = Z.Rn, F = Z.F;
    const double dm = (double)m;
    S.EmL = Z.cA * (lc(0, 0, m) - Z.rho_lt * lc(1, 0, m - lt));
    S.EpL = Z.cA_rhoinv_ltm1 * (la(0, m - lt + 1) - Z.rho_lt * la(1, m + 1));
    S.EpR = Z.cA_rhoinv_Rn * (lc(2, 0, m - L + Rn) - Z.rho_Rn * lc(3, 0, m - L));
    S.EmR = Z.cA_rho * (la(2, m - L + 1) - Z.rho_Rn * la(3, m - L + Rn + 1));
    const double a0 = d0(m) - d0(m - lt), a1 = lc(0, 1, m) - lc(1, 1, m - lt), a2 = lc(0, 2, m) - lc(1, 2, m - lt);
    S.W0L = a0;
    S.W1L = dm * a0 - a1;
    S.W2L = dm * dm * a0 - 2.0 * dm * a1 + a2;
    S.W0F = d0(m - lt) - d0(m - lt - F - 1);
    const double b0 = d0(m - L + Rn) - d0(m - L);
    const double b1 = lc(2, 1, m - L + Rn) - lc(3, 1, m - L), b2 = lc(2, 2, m - L + Rn) - lc(3, 2, m - L);
    const double qq = (double)(m - L);
    S.V0 = b0;
    S.V1 = b1 - qq * b0;
    S.V2 = b2 - 2.0 * qq * b1 + qq * qq * b0;
}

// one input stream d[j], j advancing by one per step
struct CzStream {
    const double* p;   // &TT[j+1] of the NEXT d to produce (fast path) / TT base (safe path)
    double t, yprev;
    int j;
    __device__ __forceinline__ void init(const double* TT, int j0)
    {
        j = j0;
        t = TT[max(j0, 0)];
        yprev = (j0 >= 1) ? t - TT[j0 - 1] : 0.0;
        p = TT + j0 + 1;
    }
    // all indices known to be inside [1, n]: no clamping
    __device__ __forceinline__ double next_fast(double r, int k)
    {
        const double tn = p[k];
        const double y = tn - t;
        const double d = fma(-r, yprev, y);
        yprev = y;
        t = tn;
        return d;
    }
    // indices before the trace read as zero, beyond the end clamp
    __device__ __forceinline__ double next_safe(const double* TT, double r, int n)
    {
        const int jj = min(max(j + 1, 0), n);
        const double tn = TT[jj];
        const double y = tn - t;
        const double d = fma(-r, yprev, y);
        yprev = y;
        t = tn;
        ++j;
        return d;
    }
};

// CH recurrence steps; emits CUSP and/or ZAC outputs, tracks (max, first argmax), fills the pick-off windows
__device__ __noinline__ void cz_run(const CzDev& Z, const double* TT, int n, int tid, CzState& S, bool want_cusp, bool want_zac,
                       int from_cusp, int from_zac, int n_w, double* stash_cusp, double* stash_zac, double (&czmax)[2],
                       int (&czarg)[2])
{
    if (!S.active) return;
    const int m0 = tid * CH;
    const int L = Z.L, lt = Z.lt, F = Z.F;
    const double r = Z.r;
    CzStream s0, s1, s2, s3;
    s0.init(TT, m0 + 1);
    s1.init(TT, m0 + 1 - lt);
    s2.init(TT, m0 - lt - F);
    s3.init(TT, m0 + 1 - L);
    auto emit = [&](int m, bool stash) {
        const int j = m - L + 1;
        const double ylast = s3.yprev;   // y[m-L]
        const double Dc = (S.EpL - S.EmL + S.EpR - S.EmR) + S.W0F;
        if (want_cusp) {
            const double o = fma(Z.g, Dc, Z.gclast_cusp * ylast);
            if (o > czmax[0]) { czmax[0] = o; czarg[0] = j; }
            if (stash) {
                const int q = j - from_cusp;
                if (q >= 0 && q < n_w) stash_cusp[q] = o;
            }
        }
        if (want_zac) {
            const double poly = (S.W2L - Z.h2 * S.W1L) + (S.V2 - Z.h2 * S.V1);
            const double o = fma(Z.g, fma(Z.B, poly, Dc), Z.gclast_zac * ylast);
            if (o > czmax[1]) { czmax[1] = o; czarg[1] = j; }
            if (stash) {
                const int q = j - from_zac;
                if (q >= 0 && q < n_w) stash_zac[q] = o;
            }
        }
    };
    auto update = [&](double a, double b, double c, double d) {
        S.EmL = fma(Z.rho, S.EmL, fma(Z.cA, a, -Z.cA_rho_lt * b));
        S.EpL = fma(Z.rho_inv, S.EpL, fma(Z.cA, a, -Z.cA_rhoinv_lt * b));
        S.W2L = S.W2L + 2.0 * S.W1L + S.W0L - Z.lt2_d * b;
        S.W1L = S.W1L + S.W0L - Z.lt_d * b;
        S.W0L = S.W0L + a - b;
        S.W0F = S.W0F + b - c;
        S.V2 = S.V2 - 2.0 * S.V1 + S.V0 + Z.Rn2_d * c;
        S.V1 = S.V1 - S.V0 + Z.Rn_d * c;
        S.V0 = S.V0 + c - d;
        S.EpR = fma(Z.rho, S.EpR, fma(Z.cA_rhoinv_Rn, c, -Z.cA * d));
        S.EmR = fma(Z.rho_inv, S.EmR, fma(Z.cA_rho_Rn, c, -Z.cA * d));
    };
    // interior chunk: every stream index is inside the trace, every m is a valid output
    const bool interior = (m0 - L >= 1) && (m0 + CH + 1 <= n);
    const int jlo = m0 - L + 1, jhi = jlo + CH - 1;
    const bool touches_window = (want_cusp && jhi >= from_cusp && jlo < from_cusp + n_w) ||
                                (want_zac && jhi >= from_zac && jlo < from_zac + n_w);
    if (interior && !touches_window) {
#pragma unroll 3
        for (int k = 0; k < CH; ++k) {
            emit(m0 + k, false);
            const double a = s0.next_fast(r, k), b = s1.next_fast(r, k), c = s2.next_fast(r, k), d = s3.next_fast(r, k);
            update(a, b, c, d);
        }
    } else {
#pragma unroll 1
        for (int k = 0; k < CH; ++k) {
            const int m = m0 + k;
            if (m >= n) break;
            if (m >= L - 1) emit(m, true);
            const double a = s0.next_safe(TT, r, n), b = s1.next_safe(TT, r, n), c = s2.next_safe(TT, r, n),
                         d = s3.next_safe(TT, r, n);
            update(a, b, c, d);
        }
    }
}

// ==================================================================================================
// the fused kernel
// ==================================================================================================
__global__ void __launch_bounds__(NT, 2)
icpc_kernel(const __grid_constant__ IcpcDev P, const uint16_t* __restrict__ wf, long long n_events, long long ld,
            double* __restrict__ rows)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint16_t* xs = reinterpret_cast<uint16_t*>(smem + SM_XS);
    double* TT = reinterpret_cast<double*>(smem + SM_TT);
    uint32_t* masks = reinterpret_cast<uint32_t*>(smem + SM_MASK);
    double* red = reinterpret_cast<double*>(smem + SM_RED);
    double* stash = reinterpret_cast<double*>(smem + SM_STASH);
    double* row = reinterpret_cast<double*>(smem + SM_ROW);
    double* scr = reinterpret_cast<double*>(smem + SM_SCR);
    int* ibuf = reinterpret_cast<int*>(smem + SM_IBUF);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SM_BAR);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = P.n;
    const uint32_t wf_bytes = (uint32_t)n * 2u;
    const double t_first = P.t_first, dt = P.dt;
    const unsigned G = P.groups;
    const double* A_int = P.dni_A;                       // global, L1/L2 resident (4 KB)
    const double* A_sig = P.dni_A + LGDSP_MAX_DNI * 4;

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    long long e = blockIdx.x;
    if (tid == 0 && e < n_events) {
        mbar_expect_tx(bar, wf_bytes);
        tma_load_1d(xs, wf + e * ld, wf_bytes, bar);
    }
    uint32_t phase = 0;
    const int i0 = tid * CH;
    const int cvalid = max(0, min(CH, n - i0));

    for (; e < n_events; e += gridDim.x) {
        // ------------------------------------------------------------------------------------------
        // pass 1: raw samples
        // ------------------------------------------------------------------------------------------
        mbar_wait(bar, phase);
        phase ^= 1;
        const uint16_t* xp = xs + i0;
        uint32_t csum = 0, cq = 0, mn = 0xFFFFu, mx = 0;
#pragma unroll 3
        for (int k = 0; k < cvalid; ++k) {
            const uint32_t x = xp[k];
            csum += x;
            cq += csum;
            mn = min(mn, x);
            mx = max(mx, x);
        }
        // baseline regression sums (exact integers): only chunks that intersect the window
        unsigned long long blSS = 0;
        uint32_t blS = 0, blSK = 0;
        {
            const int ka = max(0, P.bl_from - i0), kb = min(cvalid - 1, P.bl_until - i0);
#pragma unroll 3
            for (int k = ka; k <= kb; ++k) {
                const uint32_t x = xp[k];
                blS += x;
                blSK += x * (uint32_t)k;
                blSS += (unsigned long long)(x * x);   // 65535^2 < 2^32
            }
        }
        // saturation counts: a sample can only equal low/high if the chunk's min/max says so
        int nlow = 0, nhigh = 0;
        if ((int)mn == P.sat_low || (int)mx == P.sat_high) {
            for (int k = 0; k < cvalid; ++k) {
                const int x = xp[k];
                nlow += (x == P.sat_low);
                nhigh += (x == P.sat_high);
            }
        }
        // block scan of the chunk sums -> exclusive prefix P_excl (exact, uint32)
        uint32_t incl = csum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        uint32_t* ured = reinterpret_cast<uint32_t*>(ibuf + IB_SCAN);
        double* dscan = scr + SC_DSCAN;
        if (lane == 31) ured[wid] = incl;
        // zero the masks of this event while we are at it (committed with atomicOr later)
#pragma unroll
        for (int q = 0; q < NMASK; ++q) masks[q * NWORDS + tid] = 0u;
        double blSd, blSSd, blSXd;
        {
            // sum_i i*x = i0*sum x + sum k*x
            double v[5] = {(double)blS, (double)blSS, (double)i0 * (double)blS + (double)blSK, (double)nlow, (double)nhigh};
            block_sum<5>(v, red, tid);
            blSd = v[0]; blSSd = v[1]; blSXd = v[2]; nlow = (int)v[3]; nhigh = (int)v[4];
        }
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) woff += (w < wid) ? ured[w] : 0u;
        const uint32_t P_excl = woff + incl - csum;
        // second-order scan: PP_excl = sum over previous chunks of (cvalid*P_excl_c + cq_c)   (exact in double)
        const double v2 = (double)cvalid * (double)P_excl + (double)cq;
        double incl2 = v2;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double t = __shfl_up_sync(FULL, incl2, o);
            if (lane >= o) incl2 += t;
        }
        {
            double mm[2] = {(double)mx, -(double)mn};
            if (lane == 31) dscan[wid] = incl2;
            block_max<2>(mm, red, tid);
            mx = (uint32_t)mm[0]; mn = (uint32_t)(-mm[1]);
        }
        // (block_max's trailing __syncthreads makes the warp totals visible)
        double woff2 = 0;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) woff2 += (w < wid) ? dscan[w] : 0.0;
        const double PP_excl = woff2 + incl2 - v2;

        // saturation run lengths (only when a saturated sample exists; block-uniform branch)
        int cons_low = 0, cons_high = 0;
        if (nlow + nhigh > 0) {
            Run rl = {0, 0, 0, cvalid}, rh = {0, 0, 0, cvalid};
            {
                int cl = 0, chh = 0;
                bool pl = true, ph = true;
                for (int k = 0; k < cvalid; ++k) {
                    const int x = xp[k];
                    const bool il = (x == P.sat_low), ih = (x == P.sat_high);
                    cl = il ? cl + 1 : 0;
                    chh = ih ? chh + 1 : 0;
                    rl.best = max(rl.best, cl);
                    rh.best = max(rh.best, chh);
                    if (pl && il) rl.pre = cl; else pl = false;
                    if (ph && ih) rh.pre = chh; else ph = false;
                }
                rl.suf = cl; rh.suf = chh;
            }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                Run bl_ = run_shfl_down(rl, o), bh_ = run_shfl_down(rh, o);
                if ((lane & (2 * o - 1)) == 0) { rl = run_merge(rl, bl_); rh = run_merge(rh, bh_); }
            }
            int* ired = reinterpret_cast<int*>(red);
            __syncthreads();
            if (lane == 0) {
                ired[wid * 8 + 0] = rl.pre; ired[wid * 8 + 1] = rl.suf; ired[wid * 8 + 2] = rl.best; ired[wid * 8 + 3] = rl.len;
                ired[wid * 8 + 4] = rh.pre; ired[wid * 8 + 5] = rh.suf; ired[wid * 8 + 6] = rh.best; ired[wid * 8 + 7] = rh.len;
            }
            __syncthreads();
            Run al = {ired[0], ired[1], ired[2], ired[3]}, ah = {ired[4], ired[5], ired[6], ired[7]};
            for (int w = 1; w < NWARP; ++w) {
                Run bl_ = {ired[w * 8], ired[w * 8 + 1], ired[w * 8 + 2], ired[w * 8 + 3]};
                Run bh_ = {ired[w * 8 + 4], ired[w * 8 + 5], ired[w * 8 + 6], ired[w * 8 + 7]};
                al = run_merge(al, bl_);
                ah = run_merge(ah, bh_);
            }
            cons_low = al.best; cons_high = ah.best;
            __syncthreads();
        }

        // baseline statistics (exact sums) -> blmean
        const int bl_n = P.bl_until - P.bl_from + 1;
        double bsX, bsXX;
        xsums(P.bl_from, P.bl_until, t_first, dt, bsX, bsXX);
        const Stats bl = stats_finalize(bl_n, bsX, bsXX, blSd, blSSd, t_first * blSd + dt * blSXd);
        const double m = bl.mean;
        const double e_max = (double)mx - m, e_min = (double)mn - m;
        double thr[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) thr[k] = e_max * P.tx_frac[k];

        // ------------------------------------------------------------------------------------------
        // pass 2: pole-zero waveform, prefix sums, tail statistics, t10..t99 masks
        // ------------------------------------------------------------------------------------------
        double tl_S = 0, tl_SS = 0, tl_SX = 0, pz_S = 0, pz_SS = 0, pz_SX = 0, tl_bad = 0;
        {
            unsigned long long mb[5] = {0, 0, 0, 0, 0};
            uint32_t Pr = P_excl;
            double PPr = PP_excl;
            const double km1 = P.km1;
            double ip1 = (double)i0;                                  // becomes i+1 inside the loop
            double tri = 0.5 * (double)i0 * ((double)i0 + 1.0);       // (i+1)(i+2)/2 after the update
            double* tp = TT + i0 + 1;
            const int ta = max(0, P.tail_from - i0), tb = min(cvalid - 1, P.tail_until - i0);
            const bool has_tail = ta <= tb;
            auto body = [&](int k, bool tail) {
                const uint32_t x = xp[k];
                Pr += x;
                const double Pd = (double)Pr;
                PPr += Pd;
                ip1 += 1.0;
                tri += ip1;
                const double Sd = fma(-ip1, m, Pd);                // cumsum(w)[i]
                const double w = (double)x - m;
                const double y = fma(km1, Sd, w);                   // pole-zero corrected sample
                const double SS = fma(-tri, m, PPr);                // cumsum(cumsum(w))[i]
                tp[k] = fma(km1, SS, Sd);                           // cumsum(y)[i]
#pragma unroll
                for (int t = 0; t < 5; ++t) mb[t] |= (y >= thr[t]) ? (1ull << k) : 0ull;
                if (tail && k >= ta && k <= tb) {
                    const double X = t_first + (ip1 - 1.0) * dt;
                    pz_S += y;
                    pz_SS = fma(y, y, pz_SS);
                    pz_SX = fma(X, y, pz_SX);
                }
            };
            if (!has_tail) {
                int k = 0;
#pragma unroll 1
                for (; k + 3 <= cvalid; k += 3) { body(k, false); body(k + 1, false); body(k + 2, false); }
#pragma unroll 1
                for (; k < cvalid; ++k) body(k, false);
            } else {
#pragma unroll 1
                for (int k = 0; k < cvalid; ++k) body(k, true);
            }
            if (tid == 0) TT[0] = 0.0;
            if (G & LGDSP_GROUP_TIMING) {
#pragma unroll
                for (int q = 0; q < 5; ++q) mask_commit(masks + (M_T10 + q) * NWORDS, tid, mb[q]);
            }
            // tailstats: log-regression on the PRE-PZ waveform (src/tailstats.jl:22-72)
            if (has_tail) {
#pragma unroll 1
                for (int k = ta; k <= tb; ++k) {
                    const double w = (double)xp[k] - m;
                    if (w <= 0.0) {
                        tl_bad = 1.0;
                    } else {
                        const double X = t_first + (double)(i0 + k) * dt;
                        const double lg = log(w);
                        tl_S += lg;
                        tl_SS = fma(lg, lg, tl_SS);
                        tl_SX = fma(X, lg, tl_SX);
                    }
                }
            }
        }
        {
            double v[7] = {tl_S, tl_SS, tl_SX, pz_S, pz_SS, pz_SX, tl_bad};
            block_sum<7>(v, red, tid);
            tl_S = v[0]; tl_SS = v[1]; tl_SX = v[2]; pz_S = v[3]; pz_SS = v[4]; pz_SX = v[5]; tl_bad = v[6];
        }
        // xs is free now: prefetch the next event (TMA, async proxy) -- unless the structured CUSP/ZAC pass borrows
        // xs for its prefix tables; then the prefetch is issued inside that pass
        const bool cz_structured = (G & LGDSP_GROUP_CUSPZAC) && !P.direct;
        auto prefetch_next = [&]() {
            if (tid == 0) {
                const long long en = e + gridDim.x;
                if (en < n_events) {
                    fence_proxy_async();
                    mbar_expect_tx(bar, wf_bytes);
                    tma_load_1d(xs, wf + en * ld, wf_bytes, bar);
                }
            }
        };
        if (!cz_structured) prefetch_next();

        // resolve t10..t99 now (t50 positions the energy pick-off windows)
        if (wid < 5) {
            int pos, mult;
            resolve_runs(masks + (M_T10 + wid) * NWORDS, P.tx_min_n, lane, pos, mult);
            if (lane == 0) ibuf[IB_POS0 + M_T10 + wid] = pos;
        }
        __syncthreads();
        // t50 [us] and the DNI windows of the three energy pick-offs
        double t50_us = 0.0;
        {
            const int pos = ibuf[IB_POS0 + M_T50];
            if (pos >= 1) {
                const double x = cross_x(thr[1], y_at(TT, pos - 1), y_at(TT, pos), t_first + (double)(pos - 1) * dt, dt);
                t50_us = x * 0.001;
            }
        }
        double pk_p[3];
        int pk_from[3];
        {
            const int Ls[3] = {P.etrap.L, P.cusp_L, P.zac_L};
            const double picks[3] = {P.trap_pick, P.cusp_pick, P.zac_pick};
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                const double tf = t_first + (double)(Ls[f] - 1) * dt;
                dni_window(P.sig_dni.n_w, n - Ls[f] + 1, (t50_us * 1000.0 + picks[f] - tf) / dt, pk_p[f], pk_from[f]);
            }
        }

        // ------------------------------------------------------------------------------------------
        // pass 3: trapezoids and currents
        // ------------------------------------------------------------------------------------------
        double e10410 = -CUDART_INF, e10410n = -CUDART_INF, e535 = -CUDART_INF, e313 = -CUDART_INF, e313n = -CUDART_INF;
        double etmax = -CUDART_INF;
        int etarg = 0x7fffffff;
        {
            double dummy = 0;
            int idummy = 0;
            unsigned long long b0 = 0, b0i = 0, bd = 0;
            if (G & LGDSP_GROUP_TIMING) {
                double v = -CUDART_INF;
                trap_chunk<false, false, true>(TT, P.t0, i0, P.t0_thr, v, dummy, idummy, b0, b0i);
                if (!P.t0inv_same) {
                    b0i = 0;
                    trap_chunk<false, false, true>(TT, P.t0inv, i0, P.t0_thr, v, dummy, idummy, bd, b0i);
                }
                mask_commit(masks + M_T0 * NWORDS, tid, b0);
                mask_commit(masks + M_T0INV * NWORDS, tid, b0i);
            }
            if (G & LGDSP_GROUP_TRAPS) {
                trap_chunk<true, false, false>(TT, P.e10410, i0, 0.0, e10410, e10410n, idummy, bd, bd);
                trap_chunk<false, false, false>(TT, P.e535, i0, 0.0, e535, dummy, idummy, bd, bd);
                trap_chunk<true, false, false>(TT, P.e313, i0, 0.0, e313, e313n, idummy, bd, bd);
                trap_chunk<false, true, false>(TT, P.etrap, i0, 0.0, etmax, dummy, etarg, bd, bd);
            }
        }
        // currents: windowed first-argmax of the three SG traces and of the derivative; sg[0] full-trace stats
        double cmax[4] = {-CUDART_INF, -CUDART_INF, -CUDART_INF, -CUDART_INF};
        int carg[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
        double sg_max = -CUDART_INF, sg_S = 0, sg_SS = 0;
        const int nsg = P.sg[0].nout;
        if (G & LGDSP_GROUP_CURRENT) {
            {
                const int cnt = min(CH, nsg - i0);
                const int wa = P.cur_from[0] - i0, wb = P.cur_until[0] - i0;
                const int sa = P.intr_from - i0, sb = P.intr_until - i0;
                const bool plain = (wb < 0 || wa >= CH) && (sb < 0 || sa >= CH);   // no window touches this chunk
                if (plain) {
                    sg_chunk(TT, P.sg[0], i0, cnt, [&](int k, double s) { sg_max = fmax(sg_max, s); });
                } else {
                    sg_chunk(TT, P.sg[0], i0, cnt, [&](int k, double s) {
                        sg_max = fmax(sg_max, s);
                        if (k >= wa && k <= wb && s > cmax[0]) { cmax[0] = s; carg[0] = i0 + k; }
                        if (k >= sa && k <= sb) { sg_S += s; sg_SS = fma(s, s, sg_SS); }
                    });
                }
            }
#pragma unroll
            for (int f = 1; f < 3; ++f) {
                if (P.sg_alias[f] >= 0) continue;   // identical to an earlier filter: copied after the reduction
                const int wa = P.cur_from[f] - i0, wb = P.cur_until[f] - i0;
                if (wb >= 0 && wa < CH) {
                    const int cnt = min(CH, P.sg[f].nout - i0);
                    sg_chunk(TT, P.sg[f], i0, cnt, [&](int k, double s) {
                        if (k >= wa && k <= wb && s > cmax[f]) { cmax[f] = s; carg[f] = i0 + k; }
                    });
                }
            }
            {
                const int ka = max(0, P.cur_from[3] - i0), kb = min(cvalid - 1, P.cur_until[3] - i0);
                for (int k = ka; k <= kb; ++k) {
                    const double d = deriv_at(TT, i0 + k);
                    if (d > cmax[3]) { cmax[3] = d; carg[3] = i0 + k; }
                }
            }
        }
        // CUSP / ZAC, direct form (validation mode only)
        double czmax[2] = {-CUDART_INF, -CUDART_INF};
        int czarg[2] = {0x7fffffff, 0x7fffffff};
        if ((G & LGDSP_GROUP_CUSPZAC) && P.direct) {
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                const int L = f ? P.zac_L : P.cusp_L;
                const double* g = f ? P.zac_g : P.cusp_g;
                const int nout = n - L + 1;
                for (int j = tid; j < nout; j += NT) {
                    const double o = fir_at(TT, g, L, j);
                    if (o > czmax[f]) { czmax[f] = o; czarg[f] = j; }
                    const int r = j - pk_from[1 + f];
                    if (r >= 0 && r < P.sig_dni.n_w) stash[(1 + f) * LGDSP_MAX_DNI + r] = o;
                }
            }
        }
        // reductions of pass 3
        {
            double v[6] = {e10410, e10410n, e535, e313, e313n, sg_max};
            block_max<6>(v, red, tid);
            e10410 = v[0]; e10410n = v[1]; e535 = v[2]; e313 = v[3]; e313n = v[4]; sg_max = v[5];
        }
        {
            double v[2] = {sg_S, sg_SS};
            block_sum<2>(v, red, tid);
            sg_S = v[0]; sg_SS = v[1];
        }

        // ------------------------------------------------------------------------------------------
        // pass 4: masks on the sg[0] trace (t50_current, in-trace pile-up on the REVERSED trace)
        // ------------------------------------------------------------------------------------------
        double pile_thr = 0.0;
        const double cur_thr = sg_max * 0.5;
        if (G & LGDSP_GROUP_CURRENT) {
            const int cnt = P.intr_until - P.intr_from + 1;
            double dX, dXX;
            xsums(P.intr_from, P.intr_until, t_first + P.sg[0].offset * dt, dt, dX, dXX);
            const Stats st = stats_finalize(cnt, dX, dXX, sg_S, sg_SS, 0.0);
            pile_thr = st.sigma * P.nsigma;
            if (pile_thr == 0.0) pile_thr = 1.0;  // src/dsp_routines.jl:77
            unsigned long long bc = 0, bp = 0;
            sg_chunk(TT, P.sg[0], i0, min(CH, nsg - i0), [&](int k, double s) {
                bc |= (s >= cur_thr) ? (1ull << k) : 0ull;
                bp |= (s >= pile_thr) ? (1ull << k) : 0ull;
            });
            mask_commit(masks + M_CUR * NWORDS, tid, bc);
            mask_commit_reversed(masks + M_PILE * NWORDS, tid, bp, nsg);
        }
        // the 44 trap(rt,ft) outputs of the e_trap pick-off window
        if ((G & LGDSP_GROUP_TRAPS) && tid < P.sig_dni.n_w) stash[tid] = trap_at(TT, P.etrap, pk_from[0] + tid);

        // ------------------------------------------------------------------------------------------
        // CUSP / ZAC through their analytic structure
        // ------------------------------------------------------------------------------------------
        if (cz_structured) {
            double* tabA = reinterpret_cast<double*>(smem + SM_XS);
            double* tabB = reinterpret_cast<double*>(smem + SM_TAB);
            const int npass = P.cz_shared ? 1 : 2;
            for (int ps = 0; ps < npass; ++ps) {
                const CzDev& Z = P.cz[ps];
                const bool want_cusp = P.cz_shared || ps == 0, want_zac = P.cz_shared || ps == 1;
                if (ps > 0) __syncthreads();  // previous pass finished reading the tables
                cz_scan(Z, TT, n, tid, tabA, tabB, red, scr + SC_PP0);
                __syncthreads();
                CzState st;
                cz_init(Z, TT, n, tid, tabA, tabB, scr[SC_PP0], st);
                if (ps == npass - 1) {
                    __syncthreads();       // every thread has read its table entries: xs may be overwritten
                    prefetch_next();
                }
                cz_run(Z, TT, n, tid, st, want_cusp, want_zac, pk_from[1], pk_from[2], P.sig_dni.n_w,
                       stash + LGDSP_MAX_DNI, stash + 2 * LGDSP_MAX_DNI, czmax, czarg);
            }
        }
        {
            double v[7] = {etmax, cmax[0], cmax[1], cmax[2], cmax[3], czmax[0], czmax[1]};
            int ix[7] = {etarg, carg[0], carg[1], carg[2], carg[3], czarg[0], czarg[1]};
            block_argmax<7>(v, ix, red, tid);   // (its barriers also publish the masks and the stash)
            etmax = v[0]; etarg = ix[0];
#pragma unroll
            for (int f = 0; f < 4; ++f) { cmax[f] = v[1 + f]; carg[f] = ix[1 + f]; }
            czmax[0] = v[5]; czarg[0] = ix[5]; czmax[1] = v[6]; czarg[1] = ix[6];
#pragma unroll
            for (int f = 1; f < 3; ++f)
                if (P.sg_alias[f] >= 0) { cmax[f] = cmax[P.sg_alias[f]]; carg[f] = carg[P.sg_alias[f]]; }
        }
        // crossing resolution: t0, t0_inv, t50_current, pile-up (one warp each)
        if (wid < 4) {
            const int which[4] = {M_T0, M_T0INV, M_CUR, M_PILE};
            const int ks[4] = {P.t0_min_n, P.t0_min_n, P.tx_min_n, P.intr_min_n};
            int pos, mult;
            resolve_runs(masks + which[wid] * NWORDS, ks[wid], lane, pos, mult);
            if (lane == 0) {
                ibuf[IB_POS0 + which[wid]] = pos;
                if (which[wid] == M_PILE) ibuf[IB_MULT] = mult;
            }
        }
        if (tid < 64) row[tid] = 0.0;
        __syncthreads();

        // ------------------------------------------------------------------------------------------
        // pass 5: scalar results, spread over the warps; stage A
        // ------------------------------------------------------------------------------------------
        if (wid == 0) {
            if (lane == 0) {
                row[LGDSP_COL_blmean] = bl.mean; row[LGDSP_COL_blsigma] = bl.sigma;
                row[LGDSP_COL_blslope] = bl.slope; row[LGDSP_COL_bloffset] = bl.offset;
                row[LGDSP_COL_qc_label] = -1.0;
                row[LGDSP_COL_e_max] = e_max; row[LGDSP_COL_e_min] = e_min;
                row[LGDSP_COL_n_sat_low] = (double)nlow; row[LGDSP_COL_n_sat_high] = (double)nhigh;
                row[LGDSP_COL_n_sat_low_cons] = (double)cons_low; row[LGDSP_COL_n_sat_high_cons] = (double)cons_high;
            } else if (lane == 1 || lane == 2) {
                const int tn = P.tail_until - P.tail_from + 1;
                double tsX, tsXX;
                xsums(P.tail_from, P.tail_until, t_first, dt, tsX, tsXX);
                if (lane == 1) {
                    // tailstats  src/tailstats.jl:22-72
                    if (tl_bad == 0.0) {
                        const Stats ts = stats_finalize(tn, tsX, tsXX, tl_S, tl_SS, tl_SX);
                        row[LGDSP_COL_tail_mean] = ts.mean; row[LGDSP_COL_tail_sigma] = ts.sigma;
                        row[LGDSP_COL_tail_tau] = div_rn(-1.0, ts.slope);
                    }
                } else {
                    const Stats pz = stats_finalize(tn, tsX, tsXX, pz_S, pz_SS, pz_SX);
                    row[LGDSP_COL_tailmean] = pz.mean; row[LGDSP_COL_tailsigma] = pz.sigma;
                    row[LGDSP_COL_tailslope] = pz.slope; row[LGDSP_COL_tailoffset] = pz.offset;
                }
            }
        } else if (wid == 1) {
            // interpolated crossings: lanes 0..4 t10..t99, lane 5 t0, lane 6 t0_inv
            if (lane < 5) {
                const int pos = ibuf[IB_POS0 + M_T10 + lane];
                const double th = lane == 0 ? thr[0] : lane == 1 ? thr[1] : lane == 2 ? thr[2] : lane == 3 ? thr[3] : thr[4];
                double t = 0.0;
                if (pos >= 1)
                    t = cross_x(th, y_at(TT, pos - 1), y_at(TT, pos), t_first + (double)(pos - 1) * dt, dt) * 0.001;
                if (t != t) t = 0.0;
                scr[SC_TX + lane] = t;
                if (G & LGDSP_GROUP_TIMING) row[LGDSP_COL_t10 + lane] = t;
            } else if (lane == 5 || lane == 6) {
                const bool inv = lane == 6;
                const TrapDev& tr = inv ? P.t0inv : P.t0;
                const int pos = ibuf[IB_POS0 + (inv ? M_T0INV : M_T0)];
                double t = 0.0;
                if (pos >= 1) {
                    const double tl = t_first + (double)(pos - 1 + tr.L - 1) * dt;
                    const double sgn = inv ? -1.0 : 1.0;
                    t = cross_x(P.t0_thr, sgn * trap_at(TT, tr, pos - 1), sgn * trap_at(TT, tr, pos), tl, dt) * 0.001;
                    if (t != t) t = 0.0;
                }
                scr[inv ? SC_T0INV : SC_T0] = t;
                if (G & LGDSP_GROUP_TIMING) row[inv ? LGDSP_COL_t0_inv : LGDSP_COL_t0] = t;
            }
        } else if (wid == 2) {
            if (G & LGDSP_GROUP_TRAPS) {
                const double v = dni_eval_warp(A_sig, P.sig_dni.n_w, P.sig_dni.m, stash, pk_p[0] - (double)pk_from[0], lane);
                if (lane == 0) {
                    row[LGDSP_COL_e_10410] = e10410; row[LGDSP_COL_e_535] = e535; row[LGDSP_COL_e_313] = e313;
                    row[LGDSP_COL_e_10410_inv] = e10410n; row[LGDSP_COL_e_313_inv] = e313n;
                    row[LGDSP_COL_e_trap_max] = etmax;
                    row[LGDSP_COL_t_trap_max] = t_first + (double)(etarg + P.etrap.L - 1) * dt;
                    row[LGDSP_COL_e_trap] = v;
                }
            }
        } else if (wid == 3 || wid == 4) {
            if (G & LGDSP_GROUP_CUSPZAC) {
                const int f = wid - 3;
                const double v = dni_eval_warp(A_sig, P.sig_dni.n_w, P.sig_dni.m, stash + (1 + f) * LGDSP_MAX_DNI,
                                               pk_p[1 + f] - (double)pk_from[1 + f], lane);
                if (lane == 0) {
                    const int L = f ? P.zac_L : P.cusp_L;
                    row[f ? LGDSP_COL_e_zac_max : LGDSP_COL_e_cusp_max] = czmax[f];
                    row[f ? LGDSP_COL_t_zac_max : LGDSP_COL_t_cusp_max] = t_first + (double)(czarg[f] + L - 1) * dt;
                    row[f ? LGDSP_COL_e_zac : LGDSP_COL_e_cusp] = v;
                }
            }
        } else if (wid == 5) {
            if (lane < 4 && (G & LGDSP_GROUP_CURRENT)) {
                // get_wvf_maximum  src/interpolation.jl:30-46: parabola only if strictly inside the window
                const int f = lane;
                const int a = carg[f];
                double v = cmax[f];
                if (a > P.cur_from[f] && a < P.cur_until[f]) {
                    const double y1 = (f < 3) ? sg_at(TT, P.sg[f], a - 1) : deriv_at(TT, a - 1);
                    const double y3 = (f < 3) ? sg_at(TT, P.sg[f], a + 1) : deriv_at(TT, a + 1);
                    v = extrema3(y1, v, y3);
                }
                row[LGDSP_COL_a_sg + f] = v;
            }
        } else if (wid == 6) {
            if (lane < 2 && (G & LGDSP_GROUP_CURRENT)) {
                const double tf = t_first + (double)P.sg[0].offset * dt;
                if (lane == 0) {
                    // t50_current  src/dsp_icpc.jl:192-195
                    const int pos = ibuf[IB_POS0 + M_CUR];
                    double t = 0.0;
                    if (pos >= 1) {
                        t = cross_x(cur_thr, sg_at(TT, P.sg[0], pos - 1), sg_at(TT, P.sg[0], pos), tf + (double)(pos - 1) * dt, dt) * 0.001;
                        if (t != t) t = 0.0;
                    }
                    row[LGDSP_COL_t50_current] = t;
                } else {
                    // in-trace pile-up  src/dsp_routines.jl:72-82 (reversed trace r[j] = s[nsg-1-j], same time axis)
                    const int pos = ibuf[IB_POS0 + M_PILE];
                    double xi = CUDART_NAN;
                    if (pos >= 1) {
                        const double yl = sg_at(TT, P.sg[0], nsg - 1 - (pos - 1)), yr = sg_at(TT, P.sg[0], nsg - 1 - pos);
                        xi = cross_x(pile_thr, yl, yr, tf + (double)(pos - 1) * dt, dt);
                    }
                    const double last_t = tf + (double)(nsg - 1) * dt;
                    row[LGDSP_COL_inTrace_intersect] = last_t - xi;
                    row[LGDSP_COL_inTrace_n] = (double)ibuf[IB_MULT];
                }
            }
        }
        __syncthreads();
        // stage B: what needs t0 / t80 / t90
        if (wid < 2) {
            if (G & LGDSP_GROUP_QDRIFT) {
                // get_qdrift  src/dsp_routines.jl:51-64; integrator trace I[i] = TT[i+1]; warp 0: qdrift @t0, warp 1: lq @t80
                const double tns = (wid == 0 ? scr[SC_T0] : scr[SC_TX + 2]) * 1000.0;
                const double first = wid == 0 ? P.qd_first : P.lq_first, last = wid == 0 ? P.qd_last : P.lq_last;
                double a[3];
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const double ts = s == 0 ? tns : (s == 1 ? tns + first : tns + last);
                    double pc;
                    int from;
                    dni_window(P.int_dni.n_w, n, (ts - t_first) / dt, pc, from);
                    a[s] = dni_eval_warp(A_int, P.int_dni.n_w, P.int_dni.m, TT + from + 1, pc - (double)from, lane);
                }
                if (lane == 0) {
                    const double area1 = a[1] - a[0], area2 = a[2] - a[1];
                    row[wid == 0 ? LGDSP_COL_qdrift : LGDSP_COL_lq] = area2 - area1;
                }
            }
        } else if (wid == 2) {
            if (lane == 0 && (G & LGDSP_GROUP_TIMING)) row[LGDSP_COL_drift_time] = (scr[SC_TX + 3] - scr[SC_T0]) * 1000.0;
        }
        __syncthreads();
        if (tid < LGDSP_NCOL) rows[e * LGDSP_NCOL + tid] = row[tid];
        // (the next iteration's first barrier orders the reuse of row/stash/masks/scr)
    }
}

// ==================================================================================================
// Trapezoid sweep kernel: dsp_trap_rt_optimization / dsp_trap_ft_optimization
// (/root/reference/src/dsp_filter_optimization.jl:102-133, 241-274).  The reference re-filters every waveform
// once per grid point and reads ONE interpolated sample of each filtered trace; here the waveform's prefix sums
// stay resident in SMEM and every (rt, ft) variant only evaluates the n_w trapezoid outputs of its
// PolynomialDNI window (4 look-ups each).  One warp per variant.
// ==================================================================================================
constexpr int SW_XS = 0;
constexpr int SW_TT = SW_XS + MAXN * 2;
constexpr int SW_MASK = SW_TT + TT_LEN * 8;
constexpr int SW_RED = SW_MASK + NWORDS * 4;
constexpr int SW_DNI = SW_RED + NWARP * RED_W * 8;
constexpr int SW_OUT = SW_DNI + LGDSP_MAX_DNI * 4 * 8;
constexpr int SW_MAXVAR = 1024;
constexpr int SW_IBUF = SW_OUT + SW_MAXVAR * 4;
constexpr int SW_DSCAN = SW_IBUF + 64 * 4;
constexpr int SW_BAR = SW_DSCAN + 8 * 8;
constexpr int SW_TOTAL = SW_BAR + 16;

__global__ void __launch_bounds__(NT, 2)
sweep_kernel(const __grid_constant__ SweepDev P, const uint16_t* __restrict__ wf, long long n_events, long long ld,
             float* __restrict__ out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint16_t* xs = reinterpret_cast<uint16_t*>(smem + SW_XS);
    double* TT = reinterpret_cast<double*>(smem + SW_TT);
    uint32_t* mask = reinterpret_cast<uint32_t*>(smem + SW_MASK);
    double* red = reinterpret_cast<double*>(smem + SW_RED);
    double* dniA = reinterpret_cast<double*>(smem + SW_DNI);
    float* obuf = reinterpret_cast<float*>(smem + SW_OUT);
    int* ibuf = reinterpret_cast<int*>(smem + SW_IBUF);
    double* dscan = reinterpret_cast<double*>(smem + SW_DSCAN);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SW_BAR);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = P.n;
    const uint32_t wf_bytes = (uint32_t)n * 2u;
    const double t_first = P.t_first, dt = P.dt;
    for (int i = tid; i < LGDSP_MAX_DNI * 4; i += NT) dniA[i] = P.dni_A[i];
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    long long e = blockIdx.x;
    if (tid == 0 && e < n_events) {
        mbar_expect_tx(bar, wf_bytes);
        tma_load_1d(xs, wf + e * ld, wf_bytes, bar);
    }
    uint32_t phase = 0;
    const int i0 = tid * CH;
    const int cvalid = max(0, min(CH, n - i0));
    for (; e < n_events; e += gridDim.x) {
        mbar_wait(bar, phase);
        phase ^= 1;
        const uint16_t* xp = xs + i0;
        uint32_t csum = 0, cq = 0, blS = 0;
        {
            const int ka = P.bl_from - i0, kb = P.bl_until - i0;
            for (int k = 0; k < cvalid; ++k) {
                const uint32_t x = xp[k];
                csum += x;
                cq += csum;
                blS += (k >= ka && k <= kb) ? x : 0u;
            }
        }
        uint32_t incl = csum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        uint32_t* ured = reinterpret_cast<uint32_t*>(ibuf + IB_SCAN);
        if (lane == 31) ured[wid] = incl;
        mask[tid] = 0u;
        double blSd;
        {
            double v[1] = {(double)blS};
            block_sum<1>(v, red, tid);
            blSd = v[0];
        }
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) woff += (w < wid) ? ured[w] : 0u;
        const uint32_t P_excl = woff + incl - csum;
        const double v2 = (double)cvalid * (double)P_excl + (double)cq;
        double incl2 = v2;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double t = __shfl_up_sync(FULL, incl2, o);
            if (lane >= o) incl2 += t;
        }
        if (lane == 31) dscan[wid] = incl2;
        __syncthreads();
        double woff2 = 0;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) woff2 += (w < wid) ? dscan[w] : 0.0;
        const double PP_excl = woff2 + incl2 - v2;
        // blmean exactly as signalstats: mean_Y = sum_Y * inv_n
        const double m = mul_rn(blSd, div_rn(1.0, (double)(P.bl_until - P.bl_from + 1)));
        double ymax = -CUDART_INF;
        {
            uint32_t Pr = P_excl;
            double PPr = PP_excl;
            const double km1 = P.km1;
            double ip1 = (double)i0;
            double tri = 0.5 * (double)i0 * ((double)i0 + 1.0);
            double* tp = TT + i0 + 1;
            auto body = [&](int k) {
                const uint32_t x = xp[k];
                Pr += x;
                const double Pd = (double)Pr;
                PPr += Pd;
                ip1 += 1.0;
                tri += ip1;
                const double Sd = fma(-ip1, m, Pd);
                const double w = (double)x - m;
                const double y = fma(km1, Sd, w);
                const double SS = fma(-tri, m, PPr);
                tp[k] = fma(km1, SS, Sd);
                ymax = fmax(ymax, y);
            };
            {
                int k = 0;
#pragma unroll 1
                for (; k + 3 <= cvalid; k += 3) { body(k); body(k + 1); body(k + 2); }
#pragma unroll 1
                for (; k < cvalid; ++k) body(k);
            }
        }
        if (tid == 0) TT[0] = 0.0;
        {
            double v[1] = {ymax};
            block_max<1>(v, red, tid);
            ymax = v[0];
        }
        if (tid == 0) {
            const long long en = e + gridDim.x;
            if (en < n_events) {
                fence_proxy_async();
                mbar_expect_tx(bar, wf_bytes);
                tma_load_1d(xs, wf + en * ld, wf_bytes, bar);
            }
        }
        // t50 on the PZ waveform at 0.5*maximum  (src/dsp_filter_optimization.jl:260)
        const double thr = ymax * 0.5;
        {
            unsigned long long b = 0;
            const double* p = TT + i0;
            for (int k = 0; k < cvalid; ++k) b |= ((p[k + 1] - p[k]) >= thr) ? (1ull << k) : 0ull;
            mask_commit(mask, tid, b);
        }
        __syncthreads();
        if (wid == 0) {
            int pos, mult;
            resolve_runs(mask, P.tx_min_n, lane, pos, mult);
            if (lane == 0) ibuf[0] = pos;
        }
        __syncthreads();
        double t50_us = 0.0;
        {
            const int pos = ibuf[0];
            if (pos >= 1) {
                t50_us = cross_x(thr, y_at(TT, pos - 1), y_at(TT, pos), t_first + (double)(pos - 1) * dt, dt) * 0.001;
                if (t50_us != t50_us) t50_us = 0.0;
            }
        }
        const int n_w = P.sig_dni.n_w, mdeg = P.sig_dni.m;
        for (int v = wid; v < P.nvar; v += NWARP) {
            const SweepVar sv = P.vars[v];
            const int nout = n - sv.t.L + 1;
            const double tf = t_first + (double)(sv.t.L - 1) * dt;
            const double t_ns = sv.mode ? t50_us * 1000.0 + sv.pick_ns : sv.pick_ns;
            double pc;
            int from;
            dni_window(n_w, nout, (t_ns - tf) / dt, pc, from);
            double c[LGDSP_MAX_DNI_DEG + 1] = {0, 0, 0, 0};
            for (int i = lane; i < n_w; i += 32) {
                const double val = trap_at(TT, sv.t, from + i);
#pragma unroll
                for (int j = 0; j <= LGDSP_MAX_DNI_DEG; ++j)
                    if (j < mdeg) c[j] = fma(dniA[i * mdeg + j], val, c[j]);
            }
#pragma unroll
            for (int j = 0; j <= LGDSP_MAX_DNI_DEG; ++j) c[j] = warp_sum(c[j]);
            if (lane == 0) {
                const double u = pc - (double)from;
                double r = c[mdeg - 1];
                for (int j = mdeg - 2; j >= 0; --j) r = r * u + c[j];
                obuf[v] = (nout >= n_w) ? (float)r : CUDART_NAN_F;
            }
        }
        __syncthreads();
        for (int v = tid; v < P.nvar; v += NT) out[e * (long long)P.nvar + v] = obuf[v];
        __syncthreads();
    }
}

cudaError_t sweep_configure(int* max_blocks_per_sm)
{
    cudaError_t err = cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_TOTAL);
    if (err != cudaSuccess) return err;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(max_blocks_per_sm, sweep_kernel, NT, SW_TOTAL);
}

void sweep_launch(const SweepDev& P, const uint16_t* d_wf, long long n_events, long long ld, float* d_out, int grid,
                  cudaStream_t stream)
{
    sweep_kernel<<<grid, NT, SW_TOTAL, stream>>>(P, d_wf, n_events, ld, d_out);
}

void icpc_launch(const IcpcDev& P, const uint16_t* d_wf, long long n_events, long long ld, double* d_rows, int grid,
                 cudaStream_t stream)
{
    icpc_kernel<<<grid, NT, SM_TOTAL, stream>>>(P, d_wf, n_events, ld, d_rows);
}

cudaError_t icpc_configure(int* max_blocks_per_sm)
{
    cudaError_t err = cudaFuncSetAttribute(icpc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
    if (err != cudaSuccess) return err;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(max_blocks_per_sm, icpc_kernel, NT, SM_TOTAL);
}

}  // namespace lgdsp
