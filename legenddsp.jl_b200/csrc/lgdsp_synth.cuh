// Synthetic ICPC waveform generator, shared by host and device code.
// Generalises make_fake_waveform (/root/reference/test/test_dsp_icpc.jl:11-32): flat baseline, linear rise,
// exponential decay with tau, plus per-event randomisation, white noise and ADC clipping.
// Counter-based RNG (Philox4x32-10): event e, sample group g draw from counter (e_lo, e_hi, g, stream), so any
// slice of the event stream can be generated independently on any GPU or on the host.
#pragma once
#include <cstdint>
#include <cmath>
#include "../../include/lgdsp_b200.h"

#if defined(__CUDACC__)
#define LGDSP_HD __host__ __device__ __forceinline__
#else
#define LGDSP_HD inline
#endif

namespace lgdsp_synth {

struct U4 { uint32_t x, y, z, w; };

LGDSP_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo)
{
    uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
}

LGDSP_HD U4 philox4x32_10(U4 ctr, uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo(M0, ctr.x, hi0, lo0);
        mulhilo(M1, ctr.z, hi1, lo1);
        U4 n;
        n.x = hi1 ^ ctr.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ ctr.w ^ k1;
        n.w = lo0;
        ctr = n;
        k0 += W0;
        k1 += W1;
    }
    return ctr;
}

LGDSP_HD double u01(uint32_t r) { return ((double)r + 0.5) * (1.0 / 4294967296.0); }

struct EventPars {
    double base, slope;   // baseline offset [ADC] and slope [ADC/sample]
    double amp, amp2;     // pulse amplitudes (amp2 = 0: no second pulse)
    int s0, rise;         // pulse start sample and rise length
    int s1, rise2;        // second pulse
    double clip;          // ADC ceiling
};

LGDSP_HD EventPars event_pars(const lgdsp_synth_params& sp, int64_t e)
{
    EventPars p;
    const int n = sp.n_samples;
    const double scale = (double)n / 8192.0;
    p.clip = 65520.0;  // = sat_high of the reference chain (2^16 - 16, src/dsp_icpc.jl:94)
    if (sp.mode == 1) {
        // the reference's noise-free fixture: baseline 1000, ramp over samples 3000..3125 (1-based), amplitude 1e4
        p.base = 1000.0; p.slope = 0.0; p.amp = 10000.0; p.amp2 = 0.0;
        p.s0 = (int)(2999 * scale + 0.5); p.rise = (int)(125 * scale + 0.5); if (p.rise < 1) p.rise = 1;
        p.s1 = 0; p.rise2 = 1;
        return p;
    }
    U4 c0 = {(uint32_t)e, (uint32_t)((uint64_t)e >> 32), 0u, 0x45564E54u};  // stream "EVNT"
    U4 r0 = philox4x32_10(c0, (uint32_t)sp.seed, (uint32_t)(sp.seed >> 32));
    c0.z = 1u;
    U4 r1 = philox4x32_10(c0, (uint32_t)sp.seed, (uint32_t)(sp.seed >> 32));
    p.base = 9000.0 + 6000.0 * u01(r0.x);
    p.slope = (2.0 * u01(r0.y) - 1.0) * 1e-3;
    p.s0 = (int)(scale * (3000.0 + (2.0 * u01(r0.z) - 1.0) * 64.0));
    p.rise = (int)(scale * (20.0 + 105.0 * u01(r0.w)));
    if (p.rise < 1) p.rise = 1;
    const double ucls = u01(r1.x);
    const double uamp = u01(r1.y);
    // 90 % log-uniform [50, 45000]; 2 % over-range (clips); 3 % empty; 5 % with an in-trace second pulse
    p.amp2 = 0.0;
    if (ucls < 0.03) {
        p.amp = 0.0;
    } else if (ucls < 0.05) {
        p.amp = 60000.0 + 40000.0 * uamp;
    } else {
        p.amp = 50.0 * exp(uamp * 6.802394763324311 /* ln(900) */);
        if (ucls >= 0.95) p.amp2 = p.amp * (0.2 + 0.8 * u01(r1.z));
    }
    p.s1 = p.s0 + (int)(scale * (300.0 + 2700.0 * u01(r1.w)));
    p.rise2 = p.rise;
    return p;
}

LGDSP_HD double pulse_shape(int i, int s0, int rise, double inv_tau)
{
    if (i < s0) return 0.0;
    if (i < s0 + rise) return (double)(i - s0) / (double)rise;
    return exp(-(double)(i - s0 - rise) * inv_tau);
}

// four consecutive samples i0 .. i0+3 (i0 multiple of 4) of event e
LGDSP_HD void sample_group(const lgdsp_synth_params& sp, const EventPars& p, int64_t e, int i0, uint16_t out[4])
{
    double g[4] = {0.0, 0.0, 0.0, 0.0};
    if (sp.mode != 1 && sp.noise_sigma > 0.0) {
        U4 c = {(uint32_t)e, (uint32_t)((uint64_t)e >> 32), (uint32_t)(i0 >> 2), 0x4E4F4953u};  // stream "NOIS"
        U4 r = philox4x32_10(c, (uint32_t)sp.seed, (uint32_t)(sp.seed >> 32));
        const double two_pi = 6.283185307179586;
        double m0 = sqrt(-2.0 * log(u01(r.x))), m1 = sqrt(-2.0 * log(u01(r.z)));
        double a0 = two_pi * u01(r.y), a1 = two_pi * u01(r.w);
        g[0] = m0 * cos(a0); g[1] = m0 * sin(a0); g[2] = m1 * cos(a1); g[3] = m1 * sin(a1);
    }
    const double inv_tau = 1.0 / sp.tau_samples;
    for (int k = 0; k < 4; ++k) {
        int i = i0 + k;
        double v = p.base + p.slope * i + p.amp * pulse_shape(i, p.s0, p.rise, inv_tau);
        if (p.amp2 != 0.0) v += p.amp2 * pulse_shape(i, p.s1, p.rise2, inv_tau);
        v += sp.noise_sigma * g[k];
        v = floor(v + 0.5);
        if (v < 0.0) v = 0.0;
        if (v > p.clip) v = p.clip;
        out[k] = (uint16_t)v;
    }
}

}  // namespace lgdsp_synth
