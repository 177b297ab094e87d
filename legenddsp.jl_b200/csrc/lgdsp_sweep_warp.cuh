// Trapezoid sweeps, ONE WARP per waveform: dsp_trap_rt_optimization / dsp_trap_ft_optimization and the (rt, ft) grid
// (/root/reference/src/dsp_filter_optimization.jl:102-133, 241-274).  Included by lgdsp_icpc.cu.
//
// sweep_kernel (one CTA per waveform) spends a quarter of its time in block barriers between phases that are too short
// to amortise them, and builds the prefix sums of the WHOLE pole-zero waveform although a trapezoid sweep only reads
// them in a window of ~2 200 samples around t50.  Here a warp owns a waveform from the first load to the last store:
//
//   baseline   sum x, sum i*x over bl_window (128-bit loads, exact integers)                                     [:250-253]
//   pass 1     whole waveform, 8 consecutive samples per lane and step (one 128-bit load, two steps ahead):
//              exact integer prefix sums P (warp scan), y = w + km1*cumsum(w) in closed form -> max(y);
//              per 8-sample group its [min y, max y] (float, rounded outwards) and P at its start -> SMEM;
//              P and PP = cumsum(P) at every 256-sample boundary -> SMEM                                          [:255-256]
//   pass 1b    threshold 0.5*max(y) [:260]: a group entirely below / above the threshold gives a 0x00 / 0xff mask byte,
//              the others (the rising edge; every group of a noise-only event) are re-evaluated sample by sample;
//              bit-parallel Intersect run detection (resolve_runs) -> first crossing sample
//   pass 2     prefix sums TT of the pole-zero waveform ONLY in a window [lo, lo + W) around the crossing (W from the host:
//              what the variant set can reach), 9 consecutive samples per lane and step (odd stride: conflict-free 64-bit
//              shared-memory stores), closed form from P / PP exactly as sweep_kernel -> identical TT values
//   t50        linear interpolation of the crossing [:260]
//   variants   one LANE per (rt, ft) point walks the n_w outputs of its PolynomialDNI pick-off window [:262-268]
//              (4 look-ups in TT per output, fit matrix from the constant bank); results straight to global memory.
//              A variant whose window leaves [lo, lo + W) -- clamped pick-offs of events with t50 at the trace ends -- waits
//              for a second window built where it needs it (rare; the loop ends because every variant fits W on its own).
//
// No block barrier, no atomics; the warps of a CTA share nothing.  Outputs are bit-identical to sweep_kernel's (same
// arithmetic on the same exact integers), which stays the path of FIR / Savitzky-Golay variants, 32-bit samples and
// variant sets whose reach exceeds the window capacity.

constexpr int SWW_STEP = 288;        // samples per window-build step (9 per lane)
constexpr int SWW_MIN_STEPS = 6;     // the window area also holds pass 1's group table (12 KB)
constexpr int SWW_MAX_STEPS = 10;
constexpr int SWW_MAX_WARPS = 12;    // warps per CTA (launch bound)
constexpr int SWW_EXT_BYTES = 8192;  // float2 per 8-sample group
__host__ __device__ constexpr int sww_win_bytes(int steps) { return (steps * SWW_STEP + 8) * 8; }
__host__ __device__ constexpr int sww_warp_bytes(int steps) { return sww_win_bytes(steps) + NWORDS * 4 + 36 * 4 + 34 * 8; }

struct SweepDni {
    double A[LGDSP_MAX_DNI * (LGDSP_MAX_DNI_DEG + 1)];
};

__device__ __forceinline__ uint4 sww_ld8(const uint16_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// closed-form pole-zero sample (the arithmetic of sweep_kernel's P2 body)
__device__ __forceinline__ double sww_y(uint32_t x, uint32_t Pincl, double ip1, double m, double km1)
{
    const double Sd = fma(-ip1, m, u2d(Pincl));
    const double w = u2d(x) - m;
    return fma(km1, Sd, w);
}

__global__ void __launch_bounds__(SWW_MAX_WARPS * 32, 1)
sweep_warp_kernel(const __grid_constant__ SweepDev P, const __grid_constant__ SweepDni D, const uint16_t* __restrict__ wf,
                  long long n_events, long long ld, const double* __restrict__ bl_ext, void* __restrict__ out,
                  double* __restrict__ aux)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int steps = P.win_steps;
    const int W = steps * SWW_STEP;
    unsigned char* base = smem + (size_t)wib * sww_warp_bytes(steps);
    double* win = reinterpret_cast<double*>(base);
    float2* ext = reinterpret_cast<float2*>(base);
    uint32_t* pst = reinterpret_cast<uint32_t*>(base + SWW_EXT_BYTES);
    uint32_t* mask = reinterpret_cast<uint32_t*>(base + sww_win_bytes(steps));
    uint32_t* cP = mask + NWORDS;
    double* cPP = reinterpret_cast<double*>(cP + 36);

    const int n = P.n, n_it = (n + 255) >> 8;
    const double t_first = P.t_first, dt = P.dt, km1 = P.km1;
    const int n_w = P.sig_dni.n_w, mdeg = P.sig_dni.m;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);

    for (long long e = (long long)blockIdx.x * (blockDim.x >> 5) + wib; e < n_events; e += wstride) {
        const uint16_t* __restrict__ x = wf + e * ld;
        if (lane == 0 && e + wstride < n_events) tma_prefetch_l2(wf + (e + wstride) * ld, (uint32_t)n * 2u);

        // ---- baseline window: sum x, sum i*x (exact) ----
        uint32_t blS = 0;
        unsigned long long blSX = 0;
        for (int g = (P.bl_from >> 3) + lane; g <= (P.bl_until >> 3); g += 32) {
            const uint4 r = sww_ld8(x + 8 * g);
            const uint32_t v[8] = {r.x & 0xffffu, r.x >> 16, r.y & 0xffffu, r.y >> 16, r.z & 0xffffu, r.z >> 16, r.w & 0xffffu, r.w >> 16};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = 8 * g + k;
                const bool in = (i >= P.bl_from && i <= P.bl_until);
                blS += in ? v[k] : 0u;
                if (aux) blSX += in ? (unsigned long long)v[k] * (unsigned)i : 0ull;
            }
        }
        blS = __reduce_add_sync(FULL, blS);
        const double blSd = (double)blS;
        const double blSXd = aux ? warp_sum((double)blSX) : 0.0;
        const double m_own = mul_rn(blSd, P.bl_inv_n);
        const double m = bl_ext ? bl_ext[e] : m_own;

        // ---- pass 1: max(y), group table, boundary carries ----
        uint32_t carryP = 0;
        unsigned long long carryPP = 0;
        double ymax = -CUDART_INF;
        uint4 nx1 = (8 * lane < n) ? sww_ld8(x + 8 * lane) : zero4;
        uint4 nx2 = (256 + 8 * lane < n) ? sww_ld8(x + 256 + 8 * lane) : zero4;
#pragma unroll 1
        for (int it = 0; it < n_it; ++it) {
            const uint4 r = nx1;
            nx1 = nx2;
            {
                const int i2 = (it + 2) * 256 + 8 * lane;
                nx2 = (i2 < n) ? sww_ld8(x + i2) : zero4;
            }
            const int i0 = it * 256 + 8 * lane;
            const bool act = i0 < n;
            if (lane == 0) { cP[it] = carryP; cPP[it] = (double)carryPP; }
            const uint32_t v[8] = {r.x & 0xffffu, r.x >> 16, r.y & 0xffffu, r.y >> 16, r.z & 0xffffu, r.z >> 16, r.w & 0xffffu, r.w >> 16};
            uint32_t s[8];
            s[0] = v[0];
#pragma unroll
            for (int k = 1; k < 8; ++k) s[k] = s[k - 1] + v[k];
            uint32_t incl = s[7];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t Pst = carryP + incl - s[7];
            // sum over the 256 samples of the step of P(i): 256*carryP + 8*sum_l excl_l + sum_l sum_k s_l[k]
            const uint32_t tl = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
            const uint32_t sum_t = __reduce_add_sync(FULL, tl);
            const uint32_t sum_ex = __reduce_add_sync(FULL, (uint32_t)(31 - lane) * s[7]);
            carryPP += 256ull * carryP + 8ull * sum_ex + sum_t;
            carryP += __shfl_sync(FULL, incl, 31);
            double gmax = -CUDART_INF, gmin = CUDART_INF;
            const double ib = (double)(i0 + 1);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double y = sww_y(v[k], Pst + s[k], ib + (double)k, m, km1);
                gmax = y > gmax ? y : gmax;
                gmin = y < gmin ? y : gmin;
            }
            if (act) {
                ymax = gmax > ymax ? gmax : ymax;
                ext[it * 32 + lane] = make_float2(__double2float_ru(gmax), __double2float_rd(gmin));
                pst[it * 32 + lane] = Pst;
            }
        }
        ymax = warp_max(ymax);
        const double thr = ymax * 0.5;

        // ---- pass 1b: threshold mask of y >= thr, first run of >= tx_min_n samples ----
#pragma unroll 1
        for (int it = 0; it < NWORDS / 8; ++it) {
            uint32_t b = 0;
            const int i0 = it * 256 + 8 * lane;
            if (i0 < n) {
                const float2 ex = ext[it * 32 + lane];
                if ((double)ex.x >= thr) {
                    if ((double)ex.y >= thr) {
                        b = 0xffu;
                    } else {
                        const uint4 r = sww_ld8(x + i0);
                        const uint32_t v[8] = {r.x & 0xffffu, r.x >> 16, r.y & 0xffffu, r.y >> 16, r.z & 0xffffu, r.z >> 16, r.w & 0xffffu, r.w >> 16};
                        uint32_t Pr = pst[it * 32 + lane];
                        const double ib = (double)(i0 + 1);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            Pr += v[k];
                            b |= (sww_y(v[k], Pr, ib + (double)k, m, km1) >= thr) ? (1u << k) : 0u;
                        }
                    }
                }
            }
            b <<= 8 * (lane & 3);
            b |= __shfl_xor_sync(FULL, b, 1);
            b |= __shfl_xor_sync(FULL, b, 2);
            if ((lane & 3) == 0) mask[it * 8 + (lane >> 2)] = b;
        }
        __syncwarp();
        int pos, mult;
        resolve_runs(mask, P.tx_min_n, lane, pos, mult);

        // ---- pass 2 + variants ----
        const int lo_max = n > W ? n - W : 0;
        int lo = (pos >= 1 && P.win_mode == 1) ? pos + P.win_rel_lo : P.win_abs_lo;
        lo = max(0, min(lo, lo_max));
        // the crossing samples pos-1 .. pos+1 are always inside the first window (t50 is read from it)
        if (pos >= 1 && (lo > pos - 1 || pos + 1 > lo + W)) lo = max(0, min(pos - (W >> 1), lo_max));
        unsigned done = 0;
        double t50_us = 0.0;
        bool first = true;
        int rounds = 0;
#pragma unroll 1
        while (true) {
            __syncwarp();
            // prefix sums at the window start from the boundary carries + the partial 256-sample step before lo
            uint32_t cp;
            double cpp;
            {
                const int bq = lo >> 8, rem = lo & 255;
                cp = cP[bq];
                cpp = cPP[bq];
                if (rem) {
                    const int j0 = 8 * lane;
                    const uint4 r = (bq * 256 + j0 < n) ? sww_ld8(x + bq * 256 + j0) : zero4;
                    const uint32_t v[8] = {r.x & 0xffffu, r.x >> 16, r.y & 0xffffu, r.y >> 16, r.z & 0xffffu, r.z >> 16, r.w & 0xffffu, r.w >> 16};
                    uint32_t dP = 0, dPP = 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int j = j0 + k;
                        dP += (j < rem) ? v[k] : 0u;
                        dPP += (j < rem) ? v[k] * (uint32_t)(rem - j) : 0u;
                    }
                    cpp += (double)rem * (double)cp + warp_sum((double)dPP);
                    cp += __reduce_add_sync(FULL, dP);
                }
            }
            if (lane == 0) {
                const double lod = (double)lo;
                const double tri = 0.5 * lod * (lod + 1.0);
                win[0] = fma(km1, fma(-tri, m, cpp), fma(-lod, m, u2d(cp)));
            }
            uint32_t q[9];
            {
                const int i0 = lo + 9 * lane;
#pragma unroll
                for (int k = 0; k < 9; ++k) q[k] = (i0 + k < n) ? (uint32_t)__ldg(x + i0 + k) : 0u;
            }
#pragma unroll 1
            for (int j = 0; j < steps; ++j) {
                uint32_t v[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) v[k] = q[k];
                const int i0 = lo + SWW_STEP * j + 9 * lane;
                if (j + 1 < steps) {
#pragma unroll
                    for (int k = 0; k < 9; ++k) q[k] = (i0 + SWW_STEP + k < n) ? (uint32_t)__ldg(x + i0 + SWW_STEP + k) : 0u;
                }
                uint32_t s[9];
                s[0] = v[0];
#pragma unroll
                for (int k = 1; k < 9; ++k) s[k] = s[k - 1] + v[k];
                uint32_t incl = s[8];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += t;
                }
                const uint32_t Pst = cp + incl - s[8];
                uint32_t tl = s[0];
#pragma unroll
                for (int k = 1; k < 9; ++k) tl += s[k];
                const double qd = fma(9.0, u2d(incl - s[8]), u2d(tl));   // lane sum of (P(i) - cp): exact (< 2^37)
                double inclq = qd;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double t = __shfl_up_sync(FULL, inclq, o);
                    if (lane >= o) inclq += t;
                }
                // PP before this lane's first sample: cpp + 9*lane*cp + (scan of the lane sums)
                double PPr = cpp + (double)(9 * lane) * u2d(cp) + (inclq - qd);
                uint32_t Pr = Pst;
                double ip1 = (double)i0;
                double tri = 0.5 * ip1 * (ip1 + 1.0);
                double* tp = win + SWW_STEP * j + 9 * lane + 1;
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    Pr += v[k];
                    const double Pd = u2d(Pr);
                    PPr += Pd;
                    ip1 += 1.0;
                    tri += ip1;
                    const double Sd = fma(-ip1, m, Pd);
                    const double SS = fma(-tri, m, PPr);
                    tp[k] = (i0 + k < n) ? fma(km1, SS, Sd) : 0.0;
                }
                cpp += (double)SWW_STEP * u2d(cp) + __shfl_sync(FULL, inclq, 31);
                cp += __shfl_sync(FULL, incl, 31);
            }
            __syncwarp();
            const double* TTw = win - lo;   // TTw[i] = TT[i] for lo <= i <= lo + W
            if (first) {
                first = false;
                if (pos >= 1) {
                    t50_us = cross_x(thr, y_at(TTw, pos - 1), y_at(TTw, pos), t_first + (double)(pos - 1) * dt, dt) * 0.001;
                    if (t50_us != t50_us) t50_us = 0.0;
                }
                if (aux && lane == 0) {
                    const Stats st = stats_finalize(P.bl_inv_n, P.bl_sX, P.bl_sXX, blSd, 0.0, t_first * blSd + dt * blSXd);
                    double* a = aux + e * 4;
                    a[0] = m_own; a[1] = st.slope; a[2] = t50_us; a[3] = 0.0;
                }
            }
            int minfrom = 0x7fffffff;
            int rnd = 0;
#pragma unroll 1
            for (int v = lane; v < P.nvar; v += 32, ++rnd) {
                if ((done >> rnd) & 1u) continue;
                const SweepVar& sv = P.vars[v];
                const TrapDev t = sv.t;
                const double pick = sv.pick_ns;
                const int nout = n - t.L + 1;
                const double tf = __fma_rn((double)(t.L - 1), dt, t_first);
                const double t_ns = sv.mode ? __fma_rn(t50_us, 1000.0, pick) : pick;
                double pc;
                int from;
                dni_window(n_w, nout, (t_ns - tf) / dt, pc, from);
                if (from < lo || from + t.L + n_w - 1 > lo + W) {   // look-ups TT[from .. from + L + n_w - 1]
                    minfrom = min(minfrom, from);
                    continue;
                }
                done |= 1u << rnd;
                const double* p0 = TTw + from;
                const double* p1 = p0 + t.a;
                const double* p2 = p1 + t.g;
                const double* p3 = p0 + t.L;
                double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll 4
                for (int i = 0; i < n_w; ++i) {
                    const double val = __fma_rn(p3[i] - p2[i], t.inv2, -__dmul_rn(p1[i] - p0[i], t.inv1));
                    const double* a = D.A + i * mdeg;
                    c0 = fma(a[0], val, c0);
                    if (mdeg > 1) c1 = fma(a[1], val, c1);
                    if (mdeg > 2) c2 = fma(a[2], val, c2);
                    if (mdeg > 3) c3 = fma(a[3], val, c3);
                }
                const double u = pc - (double)from;
                const double res = (nout >= n_w) ? fma(fma(fma(c3, u, c2), u, c1), u, c0) : CUDART_NAN;
                if (P.out_f64) reinterpret_cast<double*>(out)[e * (long long)P.nvar + v] = res;
                else reinterpret_cast<float*>(out)[e * (long long)P.nvar + v] = (float)res;
            }
            minfrom = __reduce_min_sync(FULL, minfrom);
            if (minfrom == 0x7fffffff) break;
            if (++rounds > 2 * 1024 / 32 + 2) {
                // cannot happen (every variant fits a window that starts at its own `from`); never spin on the device
                rnd = 0;
                for (int v = lane; v < P.nvar; v += 32, ++rnd) {
                    if ((done >> rnd) & 1u) continue;
                    if (P.out_f64) reinterpret_cast<double*>(out)[e * (long long)P.nvar + v] = CUDART_NAN;
                    else reinterpret_cast<float*>(out)[e * (long long)P.nvar + v] = CUDART_NAN_F;
                }
                break;
            }
            lo = max(0, min(minfrom, lo_max));
        }
        __syncwarp();   // the window area becomes the next event's group table
    }
}

// launch geometry for a window of `steps` steps: warps per CTA and CTAs per SM that keep the most warps resident
struct SwwGeom {
    int warps_per_cta = 0, ctas_per_sm = 0;
};
inline SwwGeom sww_geometry(int steps)
{
    static SwwGeom cache[SWW_MAX_STEPS + 1];
    static bool attr_set = false;
    if (steps < SWW_MIN_STEPS || steps > SWW_MAX_STEPS) return SwwGeom{};
    if (cache[steps].warps_per_cta) return cache[steps];
    if (!attr_set) {
        int dev = 0, optin = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (cudaFuncSetAttribute(sweep_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) != cudaSuccess) return SwwGeom{};
        attr_set = true;
    }
    SwwGeom best;
    int best_warps = 0;
    for (int w = 1; w <= SWW_MAX_WARPS; ++w) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sweep_warp_kernel, w * 32, (size_t)w * sww_warp_bytes(steps)) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        if (nb * w > best_warps) { best_warps = nb * w; best.warps_per_cta = w; best.ctas_per_sm = nb; }
    }
    cache[steps] = best;
    return best;
}
