/*
 * lgdsp_synth_oracle.c -- the synthetic ICPC event stream of SURVEY.md section 8d for the CPU arms.  TEST / BENCH INFRASTRUCTURE
 * (same rules as lgdsp_oracle.c).  bench.py --impl reference must not need the product library, so the CPU arm carries its
 * own statement of the generator: counter-based Philox4x32-10 (Salmon et al., SC'11; key = seed, counter = (event, group,
 * stream tag)), the event recipe of make_fake_waveform (/root/reference/test/test_dsp_icpc.jl:11-32) with per-event
 * randomisation, Box-Muller noise, rounding and clipping at 65520 (= sat_high, src/dsp_icpc.jl:94).
 * tests/test_oracle_kat.py checks that it reproduces the product's host generator sample for sample.
 */
#include <math.h>
#include <stdint.h>
#include "../include/lgdsp_b200.h"

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

typedef struct { uint32_t v[4]; } ctr4;

static ctr4 philox(ctr4 c, uint32_t k0, uint32_t k1)
{
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c.v[0], p1 = (uint64_t)0xCD9E8D57u * c.v[2];
        ctr4 nx;
        nx.v[0] = (uint32_t)(p1 >> 32) ^ c.v[1] ^ k0;
        nx.v[1] = (uint32_t)p1;
        nx.v[2] = (uint32_t)(p0 >> 32) ^ c.v[3] ^ k1;
        nx.v[3] = (uint32_t)p0;
        c = nx;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}
static double unit(uint32_t r) { return ((double)r + 0.5) * (1.0 / 4294967296.0); }

typedef struct {
    double base, slope, amp, amp2, clip;
    int s0, rise, s1, rise2;
} event_t;

static event_t draw_event(const lgdsp_synth_params* sp, int64_t e)
{
    event_t p;
    const double scale = (double)sp->n_samples / 8192.0;
    p.clip = 65520.0;
    if (sp->mode == 1) {   /* the reference's noise-free fixture */
        p.base = 1000.0; p.slope = 0.0; p.amp = 10000.0; p.amp2 = 0.0;
        p.s0 = (int)(2999 * scale + 0.5);
        p.rise = (int)(125 * scale + 0.5);
        if (p.rise < 1) p.rise = 1;
        p.s1 = 0; p.rise2 = 1;
        return p;
    }
    const uint32_t k0 = (uint32_t)sp->seed, k1 = (uint32_t)(sp->seed >> 32);
    ctr4 c = {{(uint32_t)e, (uint32_t)((uint64_t)e >> 32), 0u, 0x45564E54u}};
    const ctr4 a = philox(c, k0, k1);
    c.v[2] = 1u;
    const ctr4 b = philox(c, k0, k1);
    p.base = 9000.0 + 6000.0 * unit(a.v[0]);
    p.slope = (2.0 * unit(a.v[1]) - 1.0) * 1e-3;
    p.s0 = (int)(scale * (3000.0 + (2.0 * unit(a.v[2]) - 1.0) * 64.0));
    p.rise = (int)(scale * (20.0 + 105.0 * unit(a.v[3])));
    if (p.rise < 1) p.rise = 1;
    const double cls = unit(b.v[0]), ua = unit(b.v[1]);
    p.amp2 = 0.0;
    if (cls < 0.03) p.amp = 0.0;                                   /* 3 % empty */
    else if (cls < 0.05) p.amp = 60000.0 + 40000.0 * ua;           /* 2 % over range */
    else {
        p.amp = 50.0 * exp(ua * 6.802394763324311);                /* log-uniform [50, 45000] */
        if (cls >= 0.95) p.amp2 = p.amp * (0.2 + 0.8 * unit(b.v[2]));   /* 5 % with a second pulse */
    }
    p.s1 = p.s0 + (int)(scale * (300.0 + 2700.0 * unit(b.v[3])));
    p.rise2 = p.rise;
    return p;
}

static double shape(int i, int s0, int rise, double inv_tau)
{
    if (i < s0) return 0.0;
    if (i < s0 + rise) return (double)(i - s0) / (double)rise;
    return exp(-(double)(i - s0 - rise) * inv_tau);
}

ORC_API int orc_synth_generate(const lgdsp_synth_params* sp, int64_t first_event, int64_t n_events, int64_t ld, uint16_t* wf)
{
    if (!sp || !wf || sp->n_samples < 4 || sp->n_samples % 4 || ld < sp->n_samples) return -1;
    const double inv_tau = 1.0 / sp->tau_samples;
    const uint32_t k0 = (uint32_t)sp->seed, k1 = (uint32_t)(sp->seed >> 32);
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < n_events; ++q) {
        const int64_t e = first_event + q;
        const event_t p = draw_event(sp, e);
        uint16_t* o = wf + q * ld;
        for (int i0 = 0; i0 < sp->n_samples; i0 += 4) {
            double g[4] = {0, 0, 0, 0};
            if (sp->mode != 1 && sp->noise_sigma > 0.0) {
                const ctr4 c = {{(uint32_t)e, (uint32_t)((uint64_t)e >> 32), (uint32_t)(i0 >> 2), 0x4E4F4953u}};
                const ctr4 r = philox(c, k0, k1);
                const double m0 = sqrt(-2.0 * log(unit(r.v[0]))), m1 = sqrt(-2.0 * log(unit(r.v[2])));
                const double a0 = 6.283185307179586 * unit(r.v[1]), a1 = 6.283185307179586 * unit(r.v[3]);
                g[0] = m0 * cos(a0); g[1] = m0 * sin(a0); g[2] = m1 * cos(a1); g[3] = m1 * sin(a1);
            }
            for (int k = 0; k < 4; ++k) {
                const int i = i0 + k;
                double v = p.base + p.slope * i + p.amp * shape(i, p.s0, p.rise, inv_tau);
                if (p.amp2 != 0.0) v += p.amp2 * shape(i, p.s1, p.rise2, inv_tau);
                v += sp->noise_sigma * g[k];
                v = floor(v + 0.5);
                if (v < 0.0) v = 0.0;
                if (v > p.clip) v = p.clip;
                o[i] = (uint16_t)v;
            }
        }
    }
    return 0;
}
