// C ABI of liblgdsp_b200 (include/lgdsp_b200.h): handle management, parameter validation/upload, launches.
// No CPU fallback: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "lgdsp_kernels.h"

using namespace lgdsp;

#define LGDSP_SPLIT_MAX_STREAMS 6

struct lgdsp_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    int icpc_bps = 0, sweep_bps = 0;
    std::string err;
    // device tables
    double* d_dniA = nullptr;      // [2][LGDSP_MAX_DNI*4]
    double* d_cusp_g = nullptr;    // [LGDSP_MAX_FIR+1]
    double* d_zac_g = nullptr;     // [LGDSP_MAX_FIR+1]
    double* d_sweep_dniA = nullptr;
    unsigned long long* d_phase = nullptr;   // [grid][8] phase cycle counters (debug)
    int phase_grid = 0;
    bool phase_on = false;
    SweepVar* d_vars = nullptr;
    int vars_cap = 0;
    double* d_taps = nullptr;      // differenced FIR / SG taps of the sweep variants
    size_t taps_cap = 0;
    IcpcDev icpc{};
    bool have_icpc = false;
    // second parameter slot: the windowed-waveform pass of dsp_icpc_compressed
    IcpcDev icpc_w{};
    bool have_icpc_w = false;
    double* d_dniA_w = nullptr;
    double* d_cusp_g_w = nullptr;
    double* d_zac_g_w = nullptr;
    void* d_cin[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [buffer][pre, wdw] staging of the compressed host path
    size_t cin_cap[2] = {0, 0};
    double* d_crows = nullptr;
    size_t crows_cap = 0;
    // host-path staging
    uint16_t* d_in[2] = {nullptr, nullptr};
    double* d_rows = nullptr;
    double* d_aux = nullptr;   // per-event external baselines / shifts (two chunks)
    size_t aux_cap = 0;
    int* d_win = nullptr;      // window list of lgdsp_window_stats_run
    void* d_sweep_out = nullptr;
    size_t in_cap = 0, rows_cap = 0, sweep_out_cap = 0;
    cudaStream_t s_copy = nullptr;
    cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    // pinned staging of the host paths (HostIO below): pageable caller memory goes through these double buffers
    void* h_pin[2] = {nullptr, nullptr};
    size_t pin_cap = 0;
    void* h_out[2] = {nullptr, nullptr};
    size_t out_cap = 0;
    cudaEvent_t ev_pin[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    int copy_threads = 4;
    int64_t host_chunk_direct = 8192;   // the same for raw samples copied directly from page-locked caller memory
    int64_t host_chunk = 16384;    // events per chunk of the host paths (4096 / 8192 / 16384 / 32768 measured: 16384 is the best of the four for pageable, pinned and encoded input)
    // encoded-waveform staging (decode_data on the device)
    uint8_t* d_enc[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [buffer][stream]
    size_t enc_cap[2] = {0, 0};
    long long* d_off[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    size_t off_cap = 0;
    void* d_dec[2] = {nullptr, nullptr};                               // decoded waveforms of the current chunk
    size_t dec_cap[2] = {0, 0};
    int* d_dstat = nullptr;
    size_t dstat_cap = 0;
    // timing
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    int64_t launches = 0;
    // split dsp_icpc pipeline (lgdsp_icpc_split.cuh): ring of prefix sums + per-event hand-over values, side streams
    int icpc_path = 1;             // 0: fused icpc_kernel, 1: split pipeline
    int split_bps[3] = {0, 0, 0};
    int64_t split_batch = 0;       // events per batch (0: default)
    int split_streams = 4;         // stream pairs the sub-batches go round (2 / 3 / 4: 6.35 / 6.56 / 6.67 M wf/s)
    double* d_tt = nullptr;
    double* d_saux = nullptr;
    double* d_scz = nullptr;       // candidate records of the CUSP/ZAC kernels
    int64_t split_cap = 0;         // allocated slots
    cudaStream_t s_split[LGDSP_SPLIT_MAX_STREAMS] = {};
    cudaStream_t s_cz[LGDSP_SPLIT_MAX_STREAMS] = {};   // CUSP/ZAC kernel next to the extract kernel
    cudaEvent_t ev_pre[LGDSP_SPLIT_MAX_STREAMS] = {}, ev_cz[LGDSP_SPLIT_MAX_STREAMS] = {};
    int split_par = 1;             // 1: CUSP/ZAC kernel on its own stream
    cudaEvent_t ev_fork = nullptr, ev_join[LGDSP_SPLIT_MAX_STREAMS] = {nullptr, nullptr, nullptr, nullptr};
};

static thread_local std::string g_create_err;

static int fail(lgdsp_handle* h, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_err = buf;
    return code;
}

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(h, e_ == cudaErrorMemoryAllocation ? LGDSP_ERR_OOM : LGDSP_ERR_CUDA, "%s: %s", \
                        #call, cudaGetErrorString(e_));                                                \
    } while (0)

extern "C" {

const char* lgdsp_version(void) { return "lgdsp_b200 0.1 (sm_100a)"; }

const char* lgdsp_last_error(const lgdsp_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int lgdsp_create(int device, void* stream, lgdsp_handle** out)
{
    lgdsp_handle* h = nullptr;
    if (!out) return fail(nullptr, LGDSP_ERR_INVALID_ARG, "lgdsp_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(nullptr, LGDSP_ERR_NO_DEVICE, "lgdsp_create: no CUDA device (%s); this library has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev) return fail(nullptr, LGDSP_ERR_INVALID_ARG, "lgdsp_create: device %d of %d", device, ndev);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(nullptr, LGDSP_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major < 10)
        return fail(nullptr, LGDSP_ERR_UNSUPPORTED, "lgdsp_create: device %d is sm_%d%d; this library is built for sm_100a only",
                    device, prop.major, prop.minor);
    h = new lgdsp_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    auto bail = [&](int code) { lgdsp_destroy(h); return code; };
    if ((e = cudaSetDevice(device)) != cudaSuccess) { fail(nullptr, LGDSP_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e)); return bail(LGDSP_ERR_CUDA); }
    if (stream) {
        h->stream = (cudaStream_t)stream;
    } else {
        if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
            fail(nullptr, LGDSP_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
            return bail(LGDSP_ERR_CUDA);
        }
        h->own_stream = true;
    }
    bool ok = cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i) {
        ok = ok && cudaEventCreateWithFlags(&h->ev_ready[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&h->ev_free[i], cudaEventDisableTiming) == cudaSuccess;
    }
    for (int i = 0; i < 2 && ok; ++i) {
        ok = ok && cudaEventCreateWithFlags(&h->ev_pin[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming) == cudaSuccess;
    }
    {
        unsigned hw = std::thread::hardware_concurrency();
        h->copy_threads = hw >= 16 ? 8 : (hw >= 4 ? (int)hw / 2 : 1);
        if (const char* env = getenv("LGDSP_COPY_THREADS")) h->copy_threads = atoi(env) > 0 ? atoi(env) : 1;
        if (const char* env = getenv("LGDSP_HOST_CHUNK")) h->host_chunk = atoll(env) > 0 ? atoll(env) : 16384;
        if (const char* env = getenv("LGDSP_HOST_CHUNK_DIRECT")) h->host_chunk_direct = atoll(env) > 0 ? atoll(env) : 8192;
    }
    ok = ok && cudaEventCreate(&h->ev0) == cudaSuccess && cudaEventCreate(&h->ev1) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_dniA, sizeof(double) * 2 * LGDSP_MAX_DNI * 4) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_cusp_g, sizeof(double) * (LGDSP_MAX_FIR + 1)) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_zac_g, sizeof(double) * (LGDSP_MAX_FIR + 1)) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_sweep_dniA, sizeof(double) * LGDSP_MAX_DNI * 4) == cudaSuccess;
    if (!ok) { fail(nullptr, LGDSP_ERR_CUDA, "lgdsp_create: resource allocation failed: %s", cudaGetErrorString(cudaGetLastError())); return bail(LGDSP_ERR_CUDA); }
    if ((e = icpc_configure(&h->icpc_bps)) != cudaSuccess || (e = sweep_configure(&h->sweep_bps)) != cudaSuccess) {
        fail(nullptr, LGDSP_ERR_CUDA, "kernel configuration failed: %s (is the library built for this GPU?)", cudaGetErrorString(e));
        return bail(LGDSP_ERR_CUDA);
    }
    if (h->icpc_bps < 1 || h->sweep_bps < 1) { fail(nullptr, LGDSP_ERR_CUDA, "kernels do not fit on an SM"); return bail(LGDSP_ERR_CUDA); }
    if ((e = icpc_split_configure(h->split_bps)) != cudaSuccess || h->split_bps[0] < 1 || h->split_bps[1] < 1 || h->split_bps[2] < 1) {
        fail(nullptr, LGDSP_ERR_CUDA, "split pipeline configuration failed: %s", cudaGetErrorString(e));
        return bail(LGDSP_ERR_CUDA);
    }
    for (int i = 0; i < LGDSP_SPLIT_MAX_STREAMS && ok; ++i) {
        ok = ok && cudaStreamCreateWithFlags(&h->s_split[i], cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaStreamCreateWithFlags(&h->s_cz[i], cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&h->ev_pre[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&h->ev_cz[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { fail(nullptr, LGDSP_ERR_CUDA, "lgdsp_create: stream allocation failed: %s", cudaGetErrorString(cudaGetLastError())); return bail(LGDSP_ERR_CUDA); }
    if (const char* env = getenv("LGDSP_ICPC_PATH")) h->icpc_path = (strcmp(env, "fused") == 0) ? 0 : 1;
    if (const char* env = getenv("LGDSP_SPLIT_BATCH")) h->split_batch = atoll(env);
    if (const char* env = getenv("LGDSP_SPLIT_STREAMS")) h->split_streams = atoi(env);
    if (const char* env = getenv("LGDSP_SPLIT_PAR")) h->split_par = atoi(env) != 0;
    if (const char* env = getenv("LGDSP_SPLIT_BPS")) {   // experiment knob: resident blocks per SM of {prefix, extract, cuspzac}, e.g. "2,3,1"
        int v[3] = {0, 0, 0};
        if (sscanf(env, "%d,%d,%d", &v[0], &v[1], &v[2]) == 3)
            for (int k = 0; k < 3; ++k)
                if (v[k] >= 1 && v[k] < h->split_bps[k]) h->split_bps[k] = v[k];
    }
    if (h->split_streams < 1) h->split_streams = 1;
    if (h->split_streams > LGDSP_SPLIT_MAX_STREAMS) h->split_streams = LGDSP_SPLIT_MAX_STREAMS;
    *out = h;
    return LGDSP_OK;
}

void lgdsp_destroy(lgdsp_handle* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int i = 0; i < LGDSP_SPLIT_MAX_STREAMS; ++i) {
        if (h->s_split[i]) { cudaStreamSynchronize(h->s_split[i]); cudaStreamDestroy(h->s_split[i]); }
        if (h->s_cz[i]) { cudaStreamSynchronize(h->s_cz[i]); cudaStreamDestroy(h->s_cz[i]); }
        if (h->ev_pre[i]) cudaEventDestroy(h->ev_pre[i]);
        if (h->ev_cz[i]) cudaEventDestroy(h->ev_cz[i]);
        if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    cudaFree(h->d_tt); cudaFree(h->d_saux); cudaFree(h->d_scz);
    cudaFree(h->d_phase);
    cudaFree(h->d_dniA); cudaFree(h->d_cusp_g); cudaFree(h->d_zac_g); cudaFree(h->d_sweep_dniA); cudaFree(h->d_vars); cudaFree(h->d_taps);
    cudaFree(h->d_in[0]); cudaFree(h->d_in[1]); cudaFree(h->d_rows); cudaFree(h->d_sweep_out);
    cudaFree(h->d_aux); cudaFree(h->d_win);
    cudaFree(h->d_dniA_w); cudaFree(h->d_cusp_g_w); cudaFree(h->d_zac_g_w); cudaFree(h->d_crows);
    for (int b = 0; b < 2; ++b) for (int k = 0; k < 2; ++k) cudaFree(h->d_cin[b][k]);
    for (int i = 0; i < 2; ++i) {
        if (h->ev_ready[i]) cudaEventDestroy(h->ev_ready[i]);
        if (h->ev_free[i]) cudaEventDestroy(h->ev_free[i]);
        if (h->ev_pin[i]) cudaEventDestroy(h->ev_pin[i]);
        if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
        if (h->h_pin[i]) cudaFreeHost(h->h_pin[i]);
        if (h->h_out[i]) cudaFreeHost(h->h_out[i]);
        for (int k = 0; k < 2; ++k) { cudaFree(h->d_enc[i][k]); cudaFree(h->d_off[i][k]); }
        cudaFree(h->d_dec[i]);
    }
    cudaFree(h->d_dstat);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->s_copy) cudaStreamDestroy(h->s_copy);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int64_t lgdsp_launch_count(const lgdsp_handle* h) { return h ? h->launches : 0; }

int lgdsp_synchronize(lgdsp_handle* h)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    return LGDSP_OK;
}

double lgdsp_last_kernel_ms(const lgdsp_handle* h)
{
    if (!h || !h->timed) return -1.0;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0;
    return (double)ms;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// parameter validation and conversion
// ---------------------------------------------------------------------------------------------------
static bool make_trap(const lgdsp_trap& t, int n, int min_out, TrapDev& d)
{
    if (t.navg < 1 || t.navg2 < 1 || t.ngap < 0) return false;
    d.a = t.navg; d.g = t.ngap; d.a2 = t.navg2;
    d.L = t.navg + t.ngap + t.navg2;
    d.nout = n - d.L + 1;
    d.pad_ = 0;
    d.inv1 = 1.0 / t.navg;
    d.inv2 = 1.0 / t.navg2;
    return d.nout >= min_out;
}


// analytic descriptor of a CUSP/ZAC filter for the structured device evaluation (lgdsp_icpc.cu)
static bool make_czdev(const lgdsp_cuspzac& z, const double* cusp_coeffs, const double* zac_coeffs, CzDev& D, std::string& why)
{
    typedef long double ld;
    const int L = z.n_taps, F = z.flat;
    const int lt = (L - F) / 2, Rn = L - lt - F - 1;
    if (lt < 2 || Rn < 1 || !(z.sigma > 0) || !(z.tau > 0)) { why = "flank too short for the structured evaluation"; return false; }
    const ld sg = z.sigma, x = (ld)lt / sg;
    if (x > 700.0L) { why = "lt/sigma too large"; return false; }
    const ld a = 1.0L / sinhl(x), cA = 0.5L * a, h = lt / 2.0L;
    const ld r = expl(-1.0L / (ld)z.tau), rho = expl(-1.0L / sg);
    memset(&D, 0, sizeof(D));
    D.L = L; D.F = F; D.lt = lt; D.Rn = Rn;
    auto mod = [](int v) { return ((v % CZ_CH) + CZ_CH) % CZ_CH; };
    D.oc[0] = 0; D.oc[1] = mod(-lt); D.oc[2] = mod(-lt - F - 1); D.oc[3] = mod(-L);
    D.oa[0] = mod(1 - lt); D.oa[1] = mod(1); D.oa[2] = mod(1 - L); D.oa[3] = mod(-lt - F);
    D.r = (double)r; D.rho = (double)rho; D.rho_inv = (double)(1.0L / rho); D.inv_sigma = (double)(1.0L / sg);
    D.cA = (double)cA;
    D.cA_rho_lt = (double)(cA * expl(-(ld)lt / sg));
    D.cA_rhoinv_lt = (double)(cA * expl((ld)lt / sg));
    D.cA_rhoinv_Rn = (double)(cA * expl((ld)Rn / sg));
    D.cA_rho_Rn = (double)(cA * expl(-(ld)Rn / sg));
    D.rho_lt = (double)expl(-(ld)lt / sg);
    D.rho_Rn = (double)expl(-(ld)Rn / sg);
    D.cA_rhoinv_ltm1 = (double)(cA * expl((ld)(lt - 1) / sg));
    D.cA_rho = (double)(cA * rho);
    {
        // capture events: causal table q after sample oc[q]; anti-causal table q after sample oa[q]-1
        struct Ev { int k, kind, tab; };
        Ev ev[8];
        for (int q = 0; q < 4; ++q) { ev[q] = {D.oc[q], 0, q}; ev[4 + q] = {D.oa[q] - 1, 1, q}; }
        for (int a_ = 0; a_ < 8; ++a_)
            for (int b_ = a_ + 1; b_ < 8; ++b_)
                if (ev[b_].k < ev[a_].k) { Ev t_ = ev[a_]; ev[a_] = ev[b_]; ev[b_] = t_; }
        D.n_ev = 8;
        for (int i = 0; i < 8; ++i) {
            D.ev_k[i] = ev[i].k; D.ev_kind[i] = ev[i].kind; D.ev_tab[i] = ev[i].tab;
            if (ev[i].kind == 0) {
                D.ev_pw[i] = (double)expl(-(ld)(ev[i].k + 1) / sg);
                D.ev_rinv[i] = 0.0;
            } else {
                const int o = ev[i].k + 1;
                D.ev_pw[i] = (double)expl(-(ld)(CZ_CH - o) / sg);
                D.ev_rinv[i] = (double)expl((ld)o / sg);
            }
        }
    }
    for (int s2 = 0; s2 < 5; ++s2) D.rho_ch_pow[s2] = (double)expl(-(ld)(CZ_CH << s2) / sg);
    for (int l = 0; l < 32; ++l) D.rho_lane[l] = (double)expl(-(ld)(CZ_CH * (l + 1)) / sg);
    D.rho_warp = (double)expl(-(ld)(CZ_CH * 32) / sg);
    D.h2 = (double)(2.0L * h);
    D.lt_d = lt; D.lt2_d = (double)lt * lt; D.Rn_d = Rn; D.Rn2_d = (double)Rn * Rn;
    // shape sums -> parabola amplitude of the ZAC
    ld acusp = 0, apar = 0;
    auto cusp_at = [&](int k) -> ld { return k < lt ? sinhl(k / sg) * a : (k <= lt + F ? 1.0L : sinhl((L - k) / sg) * a); };
    auto par_at = [&](int k) -> ld { return k < lt ? (k - h) * (k - h) - h * h : (k <= lt + F ? 0.0L : (L - k - h) * (L - k - h) - h * h); };
    for (int k = 0; k < L; ++k) { acusp += cusp_at(k); apar += par_at(k); }
    const ld B = apar != 0.0L ? -(acusp / apar) : 0.0L;
    D.B = (double)B;
    const ld g = (ld)z.beta / L;
    D.g = (double)g;
    D.gclast_cusp = (double)(g * r * cusp_at(L - 1));
    D.gclast_zac = (double)(g * r * (cusp_at(L - 1) + B * par_at(L - 1)));
    // Lipschitz constants of the two output traces (coarse-to-fine pruning in the kernel): total variation of the
    // zero-extended taps h[k] = g*(c[k] - r*c[k-1])
    for (int which = 0; which < 2; ++which) {
        const ld bb = which ? B : 0.0L;
        ld prev_c = 0.0L, prev_h = 0.0L, tv = 0.0L;
        for (int k = 0; k < L; ++k) {
            const ld ck = cusp_at(k) + bb * par_at(k);
            const ld hk = g * (ck - r * prev_c);
            tv += fabsl(hk - prev_h);
            prev_c = ck;
            prev_h = hk;
        }
        tv += fabsl(prev_h);
        (which ? D.lip_zac : D.lip_cusp) = (double)(tv * (1.0L + 1e-9L));
    }
    // the coefficient arrays passed through the ABI must be the ones this structure reproduces
    for (int which = 0; which < 2; ++which) {
        const double* co = which ? zac_coeffs : cusp_coeffs;
        if (!co) continue;
        double scale = 0;
        for (int k = 0; k < L; ++k) scale = fmax(scale, fabs(co[k]));
        const ld bb = which ? B : 0.0L;
        ld prev = 0.0L;
        for (int k = 0; k < L; ++k) {
            const ld ck = cusp_at(k) + bb * par_at(k);
            const double expect = (double)(g * (ck - r * prev));
            prev = ck;
            if (!(fabs(expect - co[k]) <= 1e-9 * scale)) {
                why = which ? "zac.coeffs do not match (sigma, flat, tau, n_taps, beta)" : "cusp.coeffs do not match (sigma, flat, tau, n_taps, beta)";
                return false;
            }
        }
    }
    return true;
}

static int icpc_prepare(lgdsp_handle* h, const lgdsp_icpc_params* p, int slot = 0)
{
    if (!p) return fail(h, LGDSP_ERR_INVALID_ARG, "params is NULL");
    if (slot && !h->d_dniA_w) {
        CK(cudaMalloc(&h->d_dniA_w, sizeof(double) * 2 * LGDSP_MAX_DNI * 4));
        CK(cudaMalloc(&h->d_cusp_g_w, sizeof(double) * (LGDSP_MAX_FIR + 1)));
        CK(cudaMalloc(&h->d_zac_g_w, sizeof(double) * (LGDSP_MAX_FIR + 1)));
    }
    double* const t_dniA = slot ? h->d_dniA_w : h->d_dniA;
    double* const t_cusp = slot ? h->d_cusp_g_w : h->d_cusp_g;
    double* const t_zac = slot ? h->d_zac_g_w : h->d_zac_g;
    if (p->struct_size != sizeof(lgdsp_icpc_params) || p->version != LGDSP_PARAMS_VERSION)
        return fail(h, LGDSP_ERR_INVALID_ARG, "lgdsp_icpc_params: size/version mismatch (got %u/%u, want %zu/%u)",
                    p->struct_size, p->version, sizeof(lgdsp_icpc_params), LGDSP_PARAMS_VERSION);
    const int n = p->n_samples;
    if (n < 64 || n > LGDSP_MAX_SAMPLES || n % 8 != 0)
        return fail(h, LGDSP_ERR_UNSUPPORTED, "n_samples = %d: need a multiple of 8 in [64, %d]", n, LGDSP_MAX_SAMPLES);
    if (!(p->dt_ns > 0) || !std::isfinite(p->t_first_ns)) return fail(h, LGDSP_ERR_INVALID_ARG, "bad time axis");
    auto win_ok = [&](int a, int b, int len) { return 0 <= a && a <= b && b <= len - 1; };
    // @assert firstindex(X) <= first(idxs) <= last(idxs) <= lastindex(X)   /root/reference/src/tailstats.jl:23-25
    if (!win_ok(p->bl_from, p->bl_until, n)) return fail(h, LGDSP_ERR_INVALID_ARG, "bl_window %d:%d outside the waveform", p->bl_from, p->bl_until);
    if (!win_ok(p->tail_from, p->tail_until, n)) return fail(h, LGDSP_ERR_INVALID_ARG, "tail_window %d:%d outside the waveform", p->tail_from, p->tail_until);
    IcpcDev D{};
    D.n = n;
    D.groups = p->groups | LGDSP_GROUP_BASE;
    D.t_first = p->t_first_ns;
    D.dt = p->dt_ns;
    D.sat_low = (p->sat_low >= 0 && p->sat_low <= 0x7fffffffLL) ? (int)p->sat_low : -1;
    D.sat_high = (p->sat_high >= 0 && p->sat_high <= 0x7fffffffLL) ? (int)p->sat_high : -1;
    D.bl_from = p->bl_from; D.bl_until = p->bl_until; D.tail_from = p->tail_from; D.tail_until = p->tail_until;
    D.km1 = p->pz_km1;
    D.bl_inv_n = 1.0 / (double)(p->bl_until - p->bl_from + 1);
    D.tail_inv_n = 1.0 / (double)(p->tail_until - p->tail_from + 1);
    {
        // sum_{i=a}^{b} X_i and X_i^2 with X_i = t_first + i*dt (configuration constants of the two regressions)
        auto xs = [&](int a, int b, double& sX, double& sXX) {
            typedef long double ld;
            const ld t0 = p->t_first_ns, dt = p->dt_ns, cnt = (ld)(b - a + 1);
            auto s2 = [](ld k) { return k * (k + 1.0L) * (2.0L * k + 1.0L) / 6.0L; };
            const ld si = 0.5L * (ld)(a + b) * cnt, sii = s2((ld)b) - s2((ld)a - 1.0L);
            sX = (double)(cnt * t0 + dt * si);
            sXX = (double)(cnt * t0 * t0 + 2.0L * t0 * dt * si + dt * dt * sii);
        };
        xs(p->bl_from, p->bl_until, D.bl_sX, D.bl_sXX);
        xs(p->tail_from, p->tail_until, D.tail_sX, D.tail_sXX);
    }
    const int nw_sig = p->sig_dni.n_w, nw_int = p->int_dni.n_w;
    auto dni_ok = [](const lgdsp_dni& d) { return d.degree >= 0 && d.degree <= LGDSP_MAX_DNI_DEG && d.n_w > d.degree && d.n_w <= LGDSP_MAX_DNI; };
    if (!dni_ok(p->sig_dni) || !dni_ok(p->int_dni)) return fail(h, LGDSP_ERR_UNSUPPORTED, "PolynomialDNI window/degree outside the supported range");
    if (nw_int > n) return fail(h, LGDSP_ERR_INVALID_ARG, "int_dni window longer than the waveform");
    D.int_dni.n_w = nw_int; D.int_dni.m = p->int_dni.degree + 1;
    D.sig_dni.n_w = nw_sig; D.sig_dni.m = p->sig_dni.degree + 1;
    if (!make_trap(p->t0_trap, n, 2, D.t0) || !make_trap(p->t0inv_trap, n, 2, D.t0inv))
        return fail(h, LGDSP_ERR_INVALID_ARG, "t0 trapezoid does not fit the waveform");
    if (!make_trap(p->trap_10410, n, 1, D.e10410) || !make_trap(p->trap_535, n, 1, D.e535) || !make_trap(p->trap_313, n, 1, D.e313))
        return fail(h, LGDSP_ERR_INVALID_ARG, "fixed trapezoid does not fit the waveform");
    if (!make_trap(p->trap_e, n, nw_sig, D.etrap)) return fail(h, LGDSP_ERR_INVALID_ARG, "trap(rt,ft) leaves fewer outputs than the DNI window");
    D.t0inv_same = (D.t0.a == D.t0inv.a && D.t0.g == D.t0inv.g && D.t0.a2 == D.t0inv.a2) ? 1 : 0;
    if (p->t0_min_n < 1 || p->tx_min_n < 1 || p->intrace_min_n < 1) return fail(h, LGDSP_ERR_INVALID_ARG, "min_n must be >= 1");
    D.t0_min_n = p->t0_min_n; D.tx_min_n = p->tx_min_n; D.direct = p->cuspzac_direct;
    D.t0_thr = p->t0_threshold;
    for (int i = 0; i < 5; ++i) D.tx_frac[i] = p->tx_frac[i];
    D.qd_first = p->qdrift_first_ns; D.qd_last = p->qdrift_last_ns; D.lq_first = p->lq_first_ns; D.lq_last = p->lq_last_ns;
    D.trap_pick = p->trap_pickoff_ns; D.cusp_pick = p->cusp_pickoff_ns; D.zac_pick = p->zac_pickoff_ns;
    for (int k = 0; k < 3; ++k) {
        const lgdsp_sg& s = p->sg[k];
        if (s.n_taps < 1 || s.n_taps > LGDSP_MAX_SG || s.n_taps > n || s.offset < 0 || s.offset >= s.n_taps)
            return fail(h, LGDSP_ERR_INVALID_ARG, "sg[%d]: bad tap count/offset", k);
        SgDev& d = D.sg[k];
        d.n_taps = s.n_taps; d.offset = s.offset; d.nout = n - s.n_taps + 1; d.pad_ = 0;
        d.gg[0] = -s.h[0];
        for (int i = 1; i < s.n_taps; ++i) d.gg[i] = s.h[i - 1] - s.h[i];
        d.gg[s.n_taps] = s.h[s.n_taps - 1];
        if (!win_ok(p->cur_from[k], p->cur_until[k], d.nout))
            return fail(h, LGDSP_ERR_INVALID_ARG, "current_window %d:%d outside sg[%d] trace", p->cur_from[k], p->cur_until[k], k);
    }
    if (!win_ok(p->cur_from[3], p->cur_until[3], n)) return fail(h, LGDSP_ERR_INVALID_ARG, "current_window outside the waveform");
    for (int k = 0; k < 4; ++k) { D.cur_from[k] = p->cur_from[k]; D.cur_until[k] = p->cur_until[k]; D.sg_alias[k] = -1; }
    for (int k = 1; k < 3; ++k)
        for (int q = 0; q < k && D.sg_alias[k] < 0; ++q)
            if (p->sg[k].n_taps == p->sg[q].n_taps && p->cur_from[k] == p->cur_from[q] && p->cur_until[k] == p->cur_until[q] &&
                memcmp(p->sg[k].h, p->sg[q].h, sizeof(double) * p->sg[k].n_taps) == 0)
                D.sg_alias[k] = q;
    D.nsigma = p->intrace_nsigma;
    D.intr_min_n = p->intrace_min_n;
    if (!win_ok(p->intrace_bl_from, p->intrace_bl_until, D.sg[0].nout)) return fail(h, LGDSP_ERR_INVALID_ARG, "in-trace sigma window outside the sg trace");
    D.intr_from = p->intrace_bl_from; D.intr_until = p->intrace_bl_until;
    D.intr_inv_n = 1.0 / (double)(p->intrace_bl_until - p->intrace_bl_from + 1);
    // CUSP / ZAC
    // (all validation first: the device tables of the handle are only touched once the whole parameter set is accepted, so
    // a rejected set leaves the previous one -- descriptors AND tables -- intact)
    const lgdsp_cuspzac* cz[2] = {&p->cusp, &p->zac};
    double* dst[2] = {t_cusp, t_zac};
    std::vector<double> g[2];
    for (int f = 0; f < 2; ++f) {
        const int L = cz[f]->n_taps;
        if (L < 4 || L > LGDSP_MAX_FIR || n - L + 1 < nw_sig)
            return fail(h, LGDSP_ERR_INVALID_ARG, "%s: %d taps do not fit (need >= %d outputs)", f ? "zac" : "cusp", L, nw_sig);
        // differenced taps on the prefix sum TT: out[j] = sum_{k=0}^{L} g[k] TT[j+L-k]
        const double* c = cz[f]->coeffs;
        g[f].resize(L + 1);
        g[f][0] = c[0];
        for (int k = 1; k < L; ++k) g[f][k] = c[k] - c[k - 1];
        g[f][L] = -c[L - 1];
    }
    D.cusp_L = p->cusp.n_taps; D.zac_L = p->zac.n_taps;
    if (!D.direct && (D.groups & LGDSP_GROUP_CUSPZAC)) {
        const lgdsp_cuspzac &c = p->cusp, &z = p->zac;
        D.cz_shared = (c.n_taps == z.n_taps && c.flat == z.flat && c.sigma == z.sigma && c.tau == z.tau && c.beta == z.beta) ? 1 : 0;
        std::string why;
        bool ok;
        if (D.cz_shared) {
            ok = make_czdev(c, c.coeffs, z.coeffs, D.cz[0], why);
            D.cz[1] = D.cz[0];
        } else {
            ok = make_czdev(c, c.coeffs, nullptr, D.cz[0], why) && make_czdev(z, nullptr, z.coeffs, D.cz[1], why);
        }
        if (!ok) return fail(h, LGDSP_ERR_UNSUPPORTED, "CUSP/ZAC structured evaluation: %s (set cuspzac_direct = 1)", why.c_str());
    }
    for (int f = 0; f < 2; ++f) CK(cudaMemcpyAsync(dst[f], g[f].data(), sizeof(double) * g[f].size(), cudaMemcpyHostToDevice, h->stream));
    std::vector<double> A(2 * LGDSP_MAX_DNI * 4, 0.0);
    memcpy(A.data(), p->int_dni.A, sizeof(double) * LGDSP_MAX_DNI * 4);
    memcpy(A.data() + LGDSP_MAX_DNI * 4, p->sig_dni.A, sizeof(double) * LGDSP_MAX_DNI * 4);
    CK(cudaMemcpyAsync(t_dniA, A.data(), sizeof(double) * A.size(), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    D.dni_A = t_dniA; D.cusp_g = t_cusp; D.zac_g = t_zac;
    if (slot) { h->icpc_w = D; h->have_icpc_w = true; }
    else { h->icpc = D; h->have_icpc = true; }
    return LGDSP_OK;
}

static int check_wf(lgdsp_handle* h, const void* wf, int64_t n_events, int64_t ld, int n, bool device, int sample_bytes = 2)
{
    if (sample_bytes != 2 && sample_bytes != 4) return fail(h, LGDSP_ERR_UNSUPPORTED, "sample_bytes = %d: uint16 (2) or uint32 (4) samples", sample_bytes);
    if (sample_bytes == 4 && n > LGDSP_MAX_SAMPLES / 2)
        return fail(h, LGDSP_ERR_UNSUPPORTED, "32-bit samples: n_samples = %d > %d", n, LGDSP_MAX_SAMPLES / 2);
    if (n_events < 0) return fail(h, LGDSP_ERR_INVALID_ARG, "n_events < 0");
    if (n_events == 0) return LGDSP_OK;  // empty table in, empty table out
    if (n_events > 0 && !wf) return fail(h, LGDSP_ERR_INVALID_ARG, "waveform pointer is NULL");
    if (ld < n) return fail(h, LGDSP_ERR_INVALID_ARG, "ld_samples (%lld) < n_samples (%d)", (long long)ld, n);
    if (device && (((uintptr_t)wf & 15u) != 0 || (ld * sample_bytes) % 16 != 0))
        return fail(h, LGDSP_ERR_INVALID_ARG, "device waveforms must be 16-byte aligned with rows of a multiple of 16 bytes (TMA bulk copy)");
    return LGDSP_OK;
}

// One dsp_icpc pass over a device batch, in stream order behind h->stream.  Fused path: one launch of icpc_kernel.  Split
// path: the batch is cut into sub-batches whose prefix sums (65.6 KB per event) fit the L2; sub-batches alternate between
// side streams (forked from / joined to h->stream with events) so that the tail of one kernel overlaps the next sub-batch.
static int icpc_dispatch(lgdsp_handle* h, const IcpcDev& D, const void* d_wf, int sample_bytes, int64_t ne, int64_t ld,
                         const double* d_bl, int64_t bl_stride, double bl_div, double* d_rows)
{
    if (h->icpc_path == 0 || D.phase_cycles != nullptr) {
        const long long cap = (long long)h->sm_count * h->icpc_bps;
        const int grid = (int)(ne < cap ? ne : cap);
        icpc_launch(D, d_wf, sample_bytes, ne, ld, d_bl, bl_stride, bl_div, d_rows, grid, h->stream);
        CK(cudaGetLastError());
        h->launches += 1;
        return LGDSP_OK;
    }
    const int S = h->split_streams;
    const int64_t B = h->split_batch > 0 ? h->split_batch : 4096;
    const int64_t need = B * S;
    if (need > h->split_cap) {
        CK(cudaStreamSynchronize(h->stream));
        for (int i = 0; i < LGDSP_SPLIT_MAX_STREAMS; ++i) CK(cudaStreamSynchronize(h->s_split[i]));
        cudaFree(h->d_tt); cudaFree(h->d_saux); cudaFree(h->d_scz);
        h->d_tt = nullptr; h->d_saux = nullptr; h->d_scz = nullptr; h->split_cap = 0;
        CK(cudaMalloc(&h->d_tt, sizeof(double) * (size_t)need * (size_t)icpc_split_tt_doubles()));
        CK(cudaMalloc(&h->d_saux, sizeof(double) * (size_t)need * (size_t)icpc_split_aux_doubles()));
        CK(cudaMalloc(&h->d_scz, sizeof(double) * (size_t)need * (size_t)icpc_split_cz_doubles()));
        h->split_cap = need;
    }
    const bool fork = S > 1 && ne > B;
    if (fork) {
        CK(cudaEventRecord(h->ev_fork, h->stream));
        for (int i = 0; i < S; ++i) CK(cudaStreamWaitEvent(h->s_split[i], h->ev_fork, 0));
    }
    const unsigned char* wf = static_cast<const unsigned char*>(d_wf);
    const bool cz = (D.groups & LGDSP_GROUP_CUSPZAC) != 0;
    int64_t b = 0;
    for (int64_t e0 = 0; e0 < ne; e0 += B, ++b) {
        const int64_t nb = (ne - e0) < B ? (ne - e0) : B;
        const int si = fork ? (int)(b % S) : 0;
        cudaStream_t st = fork ? h->s_split[si] : h->stream;
        icpc_split_launch_batch(D, wf + (size_t)e0 * (size_t)ld * (size_t)sample_bytes, sample_bytes, nb, ld,
                                d_bl ? d_bl + e0 * bl_stride : nullptr, bl_stride, bl_div, d_rows + e0 * LGDSP_NCOL,
                                h->d_tt + (size_t)si * (size_t)B * (size_t)icpc_split_tt_doubles(),
                                h->d_saux + (size_t)si * (size_t)B * (size_t)icpc_split_aux_doubles(),
                                h->d_scz + (size_t)si * (size_t)B * (size_t)icpc_split_cz_doubles(),
                                h->split_bps, h->sm_count, st,
                                h->split_par ? h->s_cz[si] : nullptr, h->ev_pre[si], h->ev_cz[si], nullptr);
        CK(cudaGetLastError());
        h->launches += cz ? (D.direct ? 3 : 4) : 2;
    }
    if (fork) {
        for (int i = 0; i < S; ++i) {
            CK(cudaEventRecord(h->ev_join[i], h->s_split[i]));
            CK(cudaStreamWaitEvent(h->stream, h->ev_join[i], 0));
        }
    }
    return LGDSP_OK;
}

extern "C" {

/* path: 0 = fused icpc_kernel, 1 = split pipeline (default); batch <= 0 / streams <= 0: keep the defaults */
int lgdsp_icpc_set_path(lgdsp_handle* h, int32_t path, int64_t batch, int32_t streams)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    if (path != 0 && path != 1) return fail(h, LGDSP_ERR_INVALID_ARG, "lgdsp_icpc_set_path: path %d (0 fused, 1 split)", path);
    h->icpc_path = path;
    if (batch > 0) h->split_batch = batch;
    if (streams > 0) h->split_streams = streams > LGDSP_SPLIT_MAX_STREAMS ? LGDSP_SPLIT_MAX_STREAMS : streams;
    return LGDSP_OK;
}

int lgdsp_icpc_set_params(lgdsp_handle* h, const lgdsp_icpc_params* p)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    return icpc_prepare(h, p);
}

static int icpc_run_device_impl(lgdsp_handle* h, const lgdsp_icpc_params* p, const void* d_wf, int sample_bytes,
                                const double* d_baseline, int64_t n_events, int64_t ld_samples, double* d_out_rows)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    if (p) { int rc = icpc_prepare(h, p); if (rc) return rc; }
    if (!h->have_icpc) return fail(h, LGDSP_ERR_INVALID_ARG, "no parameters set (pass params or call lgdsp_icpc_set_params)");
    int rc = check_wf(h, d_wf, n_events, ld_samples, h->icpc.n, true, sample_bytes);
    if (rc) return rc;
    if (n_events == 0) return LGDSP_OK;  // empty input -> empty table
    if (!d_out_rows) return fail(h, LGDSP_ERR_INVALID_ARG, "output pointer is NULL");
    const long long cap = (long long)h->sm_count * h->icpc_bps;
    const int grid = (int)(n_events < cap ? n_events : cap);
    IcpcDev D = h->icpc;
    D.phase_cycles = nullptr;
    if (h->phase_on) {
        if (!h->d_phase) CK(cudaMalloc(&h->d_phase, sizeof(unsigned long long) * (8 * (size_t)cap + 32 * 8)));
        CK(cudaMemsetAsync(h->d_phase, 0, sizeof(unsigned long long) * (8 * (size_t)cap + 32 * 8), h->stream));
        D.phase_cycles = h->d_phase;
        h->phase_grid = grid;
    }
    CK(cudaEventRecord(h->ev0, h->stream));
    rc = icpc_dispatch(h, D, d_wf, sample_bytes, n_events, ld_samples, d_baseline, 1, 1.0, d_out_rows);
    if (rc) return rc;
    CK(cudaEventRecord(h->ev1, h->stream));
    h->timed = true;
    return LGDSP_OK;
}

// per-kernel device times of the split pipeline on ONE batch run serially (prefix, extract, CUSP/ZAC select, CUSP/ZAC finish)
int lgdsp_icpc_profile_device(lgdsp_handle* h, const uint16_t* d_wf, int64_t n_events, int64_t ld_samples, double* d_out_rows,
                              double* ms4)
{
    if (!h || !ms4) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    if (!h->have_icpc) return fail(h, LGDSP_ERR_INVALID_ARG, "no parameters set");
    int rc = check_wf(h, d_wf, n_events, ld_samples, h->icpc.n, true, 2);
    if (rc) return rc;
    if (n_events <= 0 || !d_out_rows) return fail(h, LGDSP_ERR_INVALID_ARG, "nothing to profile");
    // scratch for the whole batch in one piece
    const int64_t keepB = h->split_batch;
    const int keepS = h->split_streams;
    h->split_batch = n_events; h->split_streams = 1;
    IcpcDev D = h->icpc;
    D.phase_cycles = nullptr;
    rc = icpc_dispatch(h, D, d_wf, 2, n_events, ld_samples, nullptr, 1, 1.0, d_out_rows);   // sizes the scratch, warms up
    h->split_batch = keepB; h->split_streams = keepS;
    if (rc) return rc;
    cudaEvent_t marks[5];
    for (int i = 0; i < 5; ++i) CK(cudaEventCreate(&marks[i]));
    icpc_split_launch_batch(D, d_wf, 2, n_events, ld_samples, nullptr, 1, 1.0, d_out_rows, h->d_tt, h->d_saux, h->d_scz, h->split_bps,
                            h->sm_count, h->stream, nullptr, h->ev_pre[0], h->ev_cz[0], marks);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    const bool cz = (D.groups & LGDSP_GROUP_CUSPZAC) != 0;
    float t;
    CK(cudaEventElapsedTime(&t, marks[0], marks[1])); ms4[0] = t;
    CK(cudaEventElapsedTime(&t, marks[1], marks[2])); ms4[1] = t;
    ms4[2] = ms4[3] = 0.0;
    if (cz) {
        CK(cudaEventElapsedTime(&t, marks[2], marks[3])); ms4[2] = t;
        if (!D.direct) { CK(cudaEventElapsedTime(&t, marks[3], marks[4])); ms4[3] = t; }
    }
    for (int i = 0; i < 5; ++i) cudaEventDestroy(marks[i]);
    h->launches += cz ? 4 : 2;
    return LGDSP_OK;
}

/* pinned host memory for callers that want the direct (unstaged) copy path of the host entry points */
int lgdsp_host_alloc(void** p, int64_t bytes)
{
    if (!p || bytes < 0) return LGDSP_ERR_INVALID_ARG;
    return cudaHostAlloc(p, (size_t)bytes, cudaHostAllocDefault) == cudaSuccess ? LGDSP_OK : LGDSP_ERR_OOM;
}
int lgdsp_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? LGDSP_OK : LGDSP_ERR_CUDA; }
int lgdsp_host_register(void* p, int64_t bytes)
{
    if (!p || bytes <= 0) return LGDSP_ERR_INVALID_ARG;
    return cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault) == cudaSuccess ? LGDSP_OK : LGDSP_ERR_CUDA;
}
int lgdsp_host_unregister(void* p) { return cudaHostUnregister(p) == cudaSuccess ? LGDSP_OK : LGDSP_ERR_CUDA; }

int lgdsp_icpc_run_device(lgdsp_handle* h, const lgdsp_icpc_params* p, const uint16_t* d_wf, int64_t n_events,
                          int64_t ld_samples, double* d_out_rows)
{
    return icpc_run_device_impl(h, p, d_wf, 2, nullptr, n_events, ld_samples, d_out_rows);
}

int lgdsp_icpc_run_ext_device(lgdsp_handle* h, const lgdsp_icpc_params* p, const void* d_wf, int32_t sample_bytes,
                              const double* d_baseline, int64_t n_events, int64_t ld_samples, double* d_out_rows)
{
    return icpc_run_device_impl(h, p, d_wf, sample_bytes, d_baseline, n_events, ld_samples, d_out_rows);
}

static int ensure_staging(lgdsp_handle* h, size_t in_bytes, size_t rows_bytes)
{
    if (in_bytes > h->in_cap) {
        for (int i = 0; i < 2; ++i) { cudaFree(h->d_in[i]); h->d_in[i] = nullptr; }
        h->in_cap = 0;
        for (int i = 0; i < 2; ++i) CK(cudaMalloc(&h->d_in[i], in_bytes));
        h->in_cap = in_bytes;
    }
    if (rows_bytes > h->rows_cap) {
        cudaFree(h->d_rows); h->d_rows = nullptr; h->rows_cap = 0;
        CK(cudaMalloc(&h->d_rows, rows_bytes));
        h->rows_cap = rows_bytes;
    }
    return LGDSP_OK;
}

static int ensure_aux(lgdsp_handle* h, size_t bytes)
{
    if (bytes > h->aux_cap) {
        cudaFree(h->d_aux); h->d_aux = nullptr; h->aux_cap = 0;
        CK(cudaMalloc(&h->d_aux, bytes));
        h->aux_cap = bytes;
    }
    return LGDSP_OK;
}

// ---------------------------------------------------------------------------------------------------
// Host <-> device traffic of the chunked host paths.  Caller memory that is already page-locked (cudaHostAlloc /
// cudaHostRegister / torch pin_memory) is copied directly; PAGEABLE caller memory (a Julia `flatview(wvfs.signal)`, a numpy
// array) is first packed into the handle's pinned double buffers by a few threads, so that the H2D copy of chunk k+1 and the
// D2H copy of chunk k-1 still run asynchronously beside the kernels of chunk k.  (A cudaMemcpyAsync straight from pageable
// memory is a blocking staged copy: the overlap the pipeline depends on would be gone.)
// ---------------------------------------------------------------------------------------------------
static bool is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// dst[r*dpitch .. +width) = src[r*spitch .. +width), rows split over a few threads
static void pack_rows(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows, int threads)
{
    const size_t total = width * rows;
    if (threads > 1 && total < ((size_t)4 << 20)) threads = 1;
    if (rows < (size_t)threads && !(dpitch == width && spitch == width)) threads = rows ? (int)rows : 1;
    auto work = [=](int t) {
        if (dpitch == width && spitch == width) {   // dense: split the bytes
            const size_t a0 = total * t / threads, a1 = total * (t + 1) / threads;
            memcpy(static_cast<char*>(dst) + a0, static_cast<const char*>(src) + a0, a1 - a0);
        } else {
            const size_t r0 = rows * t / threads, r1 = rows * (t + 1) / threads;
            for (size_t r = r0; r < r1; ++r)
                memcpy(static_cast<char*>(dst) + r * dpitch, static_cast<const char*>(src) + r * spitch, width);
        }
    };
    if (threads <= 1) { work(0); return; }
    std::vector<std::thread> th;
    th.reserve(threads - 1);
    for (int t = 1; t < threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
}

struct HostIO {
    lgdsp_handle* h;
    int b = 0;                 // double-buffer index of the current chunk
    size_t in_used = 0, out_used = 0;
    struct Pending { void* dst; const void* src; size_t bytes; };
    std::vector<Pending> pending[2];   // pinned -> caller copies that wait for ev_out[b]
    bool used_pin[2] = {false, false}, used_out[2] = {false, false};

    explicit HostIO(lgdsp_handle* hh) : h(hh) {}

    // capacity of the two pinned staging areas per chunk (grown on demand; growing synchronises)
    int reserve(size_t in_bytes, size_t out_bytes)
    {
        if (in_bytes > h->pin_cap) {
            CK(cudaStreamSynchronize(h->s_copy));
            for (int i = 0; i < 2; ++i) { if (h->h_pin[i]) cudaFreeHost(h->h_pin[i]); h->h_pin[i] = nullptr; }
            h->pin_cap = 0;
            for (int i = 0; i < 2; ++i) CK(cudaHostAlloc(&h->h_pin[i], in_bytes, cudaHostAllocDefault));
            h->pin_cap = in_bytes;
        }
        if (out_bytes > h->out_cap) {
            CK(cudaStreamSynchronize(h->stream));
            for (int i = 0; i < 2; ++i) { if (h->h_out[i]) cudaFreeHost(h->h_out[i]); h->h_out[i] = nullptr; }
            h->out_cap = 0;
            for (int i = 0; i < 2; ++i) CK(cudaHostAlloc(&h->h_out[i], out_bytes, cudaHostAllocDefault));
            h->out_cap = out_bytes;
        }
        return LGDSP_OK;
    }
    int flush(int bb)
    {
        if (used_out[bb]) { CK(cudaEventSynchronize(h->ev_out[bb])); used_out[bb] = false; }
        for (const Pending& q : pending[bb]) memcpy(q.dst, q.src, q.bytes);
        pending[bb].clear();
        return LGDSP_OK;
    }
    // chunk c starts: its staging buffers must be free again (the copies of chunk c-2 that used them are done)
    int begin_chunk(int c)
    {
        b = c & 1;
        in_used = out_used = 0;
        if (used_pin[b]) { CK(cudaEventSynchronize(h->ev_pin[b])); used_pin[b] = false; }
        return flush(b);
    }
    static size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }
    // rows of `width` bytes (source pitch `spitch`) -> dense device rows, on the copy stream
    int h2d(void* d_dst, const void* src, size_t spitch, size_t width, size_t rows, bool src_pinned)
    {
        if (width == 0 || rows == 0) return LGDSP_OK;
        if (src_pinned) {
            if (spitch == width) CK(cudaMemcpyAsync(d_dst, src, width * rows, cudaMemcpyHostToDevice, h->s_copy));
            else CK(cudaMemcpy2DAsync(d_dst, width, src, spitch, width, rows, cudaMemcpyHostToDevice, h->s_copy));
            return LGDSP_OK;
        }
        const size_t bytes = width * rows;
        if (in_used + bytes > h->pin_cap) return fail(h, LGDSP_ERR_CUDA, "internal: pinned input staging too small");
        char* stage = static_cast<char*>(h->h_pin[b]) + in_used;
        pack_rows(stage, width, src, spitch, width, rows, h->copy_threads);
        in_used = up256(in_used + bytes);
        CK(cudaMemcpyAsync(d_dst, stage, bytes, cudaMemcpyHostToDevice, h->s_copy));
        used_pin[b] = true;
        return LGDSP_OK;
    }
    // every H2D copy of the chunk is issued: the compute stream waits for them
    int inputs_ready()
    {
        CK(cudaEventRecord(h->ev_ready[b], h->s_copy));
        if (used_pin[b]) CK(cudaEventRecord(h->ev_pin[b], h->s_copy));
        CK(cudaStreamWaitEvent(h->stream, h->ev_ready[b], 0));
        return LGDSP_OK;
    }
    // device -> caller, in stream order behind the kernels of this chunk
    int d2h(void* dst, const void* d_src, size_t bytes, bool dst_pinned)
    {
        if (bytes == 0) return LGDSP_OK;
        if (dst_pinned) { CK(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, h->stream)); return LGDSP_OK; }
        if (out_used + bytes > h->out_cap) return fail(h, LGDSP_ERR_CUDA, "internal: pinned output staging too small");
        char* stage = static_cast<char*>(h->h_out[b]) + out_used;
        out_used = up256(out_used + bytes);
        CK(cudaMemcpyAsync(stage, d_src, bytes, cudaMemcpyDeviceToHost, h->stream));
        pending[b].push_back({dst, stage, bytes});
        used_out[b] = true;
        return LGDSP_OK;
    }
    int end_chunk()
    {
        if (used_out[b]) CK(cudaEventRecord(h->ev_out[b], h->stream));
        return LGDSP_OK;
    }
    int finish()
    {
        CK(cudaStreamSynchronize(h->s_copy));
        CK(cudaStreamSynchronize(h->stream));
        used_pin[0] = used_pin[1] = false;
        int rc = flush(0);
        return rc ? rc : flush(1);
    }
};

static int ensure_bytes(lgdsp_handle* h, void** p, size_t* cap, size_t bytes)
{
    if (bytes > *cap) {
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaStreamSynchronize(h->s_copy));
        cudaFree(*p); *p = nullptr; *cap = 0;
        CK(cudaMalloc(p, bytes));
        *cap = bytes;
    }
    return LGDSP_OK;
}

// One encoded stream set of a chunk -> decoded device rows: upload of the byte range and of the offsets slice (copy stream),
// decode kernel (compute stream, behind inputs_ready).  `slot` = 0 / 1: the two waveforms of dsp_icpc_compressed.
struct EncodedInput {
    int codec, sample_bytes, n_samples, shift;
    const uint8_t* enc;
    const int64_t* offsets;
    bool enc_pinned, off_pinned;
};
static int encoded_reserve(lgdsp_handle* h, const EncodedInput& in, int slot, int64_t chunk, int64_t n_events)
{
    size_t max_bytes = 0;
    for (int64_t e0 = 0; e0 < n_events; e0 += chunk) {
        const int64_t e1 = e0 + chunk < n_events ? e0 + chunk : n_events;
        const size_t nb = (size_t)(in.offsets[e1] - in.offsets[e0]);
        max_bytes = nb > max_bytes ? nb : max_bytes;
    }
    for (int bb = 0; bb < 2; ++bb) {
        size_t cap = h->enc_cap[slot];
        int rc = ensure_bytes(h, reinterpret_cast<void**>(&h->d_enc[bb][slot]), &cap, max_bytes + 16);
        if (rc) return rc;
        if (bb == 1) h->enc_cap[slot] = cap;
    }
    if ((size_t)(chunk + 1) * sizeof(long long) > h->off_cap) {
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaStreamSynchronize(h->s_copy));
        for (int bb = 0; bb < 2; ++bb)
            for (int k = 0; k < 2; ++k) {
                cudaFree(h->d_off[bb][k]); h->d_off[bb][k] = nullptr;
                CK(cudaMalloc(&h->d_off[bb][k], (size_t)(chunk + 1) * sizeof(long long)));
            }
        h->off_cap = (size_t)(chunk + 1) * sizeof(long long);
    }
    int rc = ensure_bytes(h, &h->d_dec[slot], &h->dec_cap[slot], (size_t)chunk * in.n_samples * in.sample_bytes);
    if (rc) return rc;
    return ensure_bytes(h, reinterpret_cast<void**>(&h->d_dstat), &h->dstat_cap, 2 * sizeof(int));
}
// d_dstat = {number of malformed streams, smallest event index among them} of the running call
static int decode_err_reset(lgdsp_handle* h)
{
    const int init[2] = {0, 0x7fffffff};
    CK(cudaMemcpyAsync(h->d_dstat, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return LGDSP_OK;
}
static int decode_err_check(lgdsp_handle* h)
{
    int st[2] = {0, 0};
    CK(cudaMemcpy(st, h->d_dstat, sizeof(st), cudaMemcpyDeviceToHost));
    if (st[0]) return fail(h, LGDSP_ERR_INVALID_ARG, "%d malformed encoded waveform(s), first at event %d", st[0], st[1]);
    return LGDSP_OK;
}
static size_t encoded_stage_bytes(const EncodedInput& in, int64_t chunk, int64_t n_events)
{
    size_t m = 0;
    for (int64_t e0 = 0; e0 < n_events; e0 += chunk) {
        const int64_t e1 = e0 + chunk < n_events ? e0 + chunk : n_events;
        const size_t nb = (size_t)(in.offsets[e1] - in.offsets[e0]) + (size_t)(e1 - e0 + 1) * sizeof(int64_t) + 512;
        m = nb > m ? nb : m;
    }
    return m;
}
static int encoded_upload(lgdsp_handle* h, HostIO& io, const EncodedInput& in, int slot, int64_t e0, int64_t ne)
{
    const size_t nb = (size_t)(in.offsets[e0 + ne] - in.offsets[e0]);
    int rc = io.h2d(h->d_enc[io.b][slot], in.enc + in.offsets[e0], nb, nb, nb ? 1 : 0, in.enc_pinned);
    if (rc) return rc;
    return io.h2d(h->d_off[io.b][slot], in.offsets + e0, (size_t)(ne + 1) * sizeof(int64_t), (size_t)(ne + 1) * sizeof(int64_t), 1,
                  in.off_pinned);
}
static int encoded_decode(lgdsp_handle* h, HostIO& io, const EncodedInput& in, int slot, int64_t e0, int64_t ne)
{
    long long longest = 0;   // sizes the decoder's per-warp stream buffers
    for (int64_t k = e0; k < e0 + ne; ++k) {
        const long long nb = (long long)(in.offsets[k + 1] - in.offsets[k]);
        longest = nb > longest ? nb : longest;
    }
    cudaError_t e = codec_decode_launch(in.codec, h->d_enc[io.b][slot], h->d_off[io.b][slot], (long long)in.offsets[e0], ne, in.n_samples,
                                        in.shift, h->d_dec[slot], in.sample_bytes, in.n_samples, nullptr, h->d_dstat, (int)e0, longest,
                                        h->sm_count, h->stream);
    if (e != cudaSuccess) return fail(h, LGDSP_ERR_CUDA, "decode kernel: %s", cudaGetErrorString(e));
    h->launches += 1;
    return LGDSP_OK;
}
static int check_encoded(lgdsp_handle* h, int codec, const void* enc, const int64_t* offsets, int64_t n_events, int sample_bytes)
{
    if (codec != LGDSP_CODEC_RADWARE && codec != LGDSP_CODEC_ULEB128ZZD) return fail(h, LGDSP_ERR_UNSUPPORTED, "unknown codec %d", codec);
    if (codec == LGDSP_CODEC_RADWARE && sample_bytes != 2) return fail(h, LGDSP_ERR_UNSUPPORTED, "RadwareSigcompress holds 16-bit samples");
    if (sample_bytes != 2 && sample_bytes != 4) return fail(h, LGDSP_ERR_UNSUPPORTED, "sample_bytes = %d", sample_bytes);
    if (n_events > 0 && (!enc || !offsets)) return fail(h, LGDSP_ERR_INVALID_ARG, "encoded data / offsets pointer is NULL");
    for (int64_t e = 0; e < n_events; ++e)
        if (offsets[e + 1] < offsets[e]) return fail(h, LGDSP_ERR_INVALID_ARG, "offsets must be non-decreasing (event %lld)", (long long)e);
    return LGDSP_OK;
}
// host buffers in, host rows out: chunked, H2D of chunk k+1 overlaps the kernels of chunk k.  `encin` != NULL: the input
// is an encoded waveform set (decode_data on the device), else raw samples `wf`.
static int icpc_run_host_impl(lgdsp_handle* h, const lgdsp_icpc_params* p, const void* wf, int sample_bytes,
                              const double* baseline, int64_t n_events, int64_t ld_samples, double* out_rows,
                              const EncodedInput* encin = nullptr)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    if (p) { int rc = icpc_prepare(h, p); if (rc) return rc; }
    if (!h->have_icpc) return fail(h, LGDSP_ERR_INVALID_ARG, "no parameters set");
    const int n = h->icpc.n;
    int rc;
    if (encin) {
        if (n_events < 0) return fail(h, LGDSP_ERR_INVALID_ARG, "n_events < 0");
        if (sample_bytes == 4 && n > LGDSP_MAX_SAMPLES / 2) return fail(h, LGDSP_ERR_UNSUPPORTED, "32-bit samples: n_samples = %d > %d", n, LGDSP_MAX_SAMPLES / 2);
        rc = check_encoded(h, encin->codec, encin->enc, encin->offsets, n_events, sample_bytes);
    } else {
        rc = check_wf(h, wf, n_events, ld_samples, n, false, sample_bytes);
    }
    if (rc) return rc;
    if (n_events == 0) return LGDSP_OK;
    if (!out_rows) return fail(h, LGDSP_ERR_INVALID_ARG, "output pointer is NULL");
    const size_t sb = (size_t)sample_bytes;
    const bool in_pinned = encin ? false : is_pinned(wf);
    // raw samples straight from page-locked memory are bound by the host link: smaller chunks shorten the un-overlapped first
    // copy / last kernel (3.25 vs 3.13 M wf/s); staged and encoded input prefers the larger chunk (2.27 vs 1.76 M wf/s pageable)
    const int64_t want = (in_pinned && h->host_chunk_direct < h->host_chunk) ? h->host_chunk_direct : h->host_chunk;
    const int64_t chunk = n_events < want ? n_events : want;
    const bool out_pinned = is_pinned(out_rows);
    const bool bl_pinned = baseline ? is_pinned(baseline) : true;
    HostIO io(h);
    size_t stage_in = 0;
    if (encin) stage_in = (encin->enc_pinned && encin->off_pinned) ? 0 : encoded_stage_bytes(*encin, chunk, n_events);
    else if (!in_pinned) stage_in = (size_t)chunk * n * sb + 256;
    if (baseline && !bl_pinned) stage_in += (size_t)chunk * sizeof(double) + 256;
    rc = io.reserve(stage_in, out_pinned ? 0 : (size_t)chunk * LGDSP_NCOL * sizeof(double) + 256);
    if (rc) return rc;
    rc = ensure_staging(h, encin ? 16 : (size_t)chunk * n * sb, (size_t)2 * chunk * LGDSP_NCOL * sizeof(double));
    if (rc) return rc;
    if (encin) {
        rc = encoded_reserve(h, *encin, 0, chunk, n_events);
        if (rc) return rc;
        rc = decode_err_reset(h);
        if (rc) return rc;
    }
    if (baseline) {
        rc = ensure_aux(h, (size_t)2 * chunk * sizeof(double));
        if (rc) return rc;
    }
    const unsigned char* src = static_cast<const unsigned char*>(wf);
    int c = 0;
    for (int64_t e0 = 0; e0 < n_events; e0 += chunk, ++c) {
        const int64_t ne = (n_events - e0) < chunk ? (n_events - e0) : chunk;
        rc = io.begin_chunk(c);
        if (rc) return rc;
        const int b = io.b;
        if (c >= 2) CK(cudaStreamWaitEvent(h->s_copy, h->ev_free[b], 0));   // the device buffers of chunk c-2 are consumed
        if (encin) rc = encoded_upload(h, io, *encin, 0, e0, ne);
        else rc = io.h2d(h->d_in[b], src + (size_t)e0 * ld_samples * sb, (size_t)ld_samples * sb, (size_t)n * sb, (size_t)ne, in_pinned);
        if (rc) return rc;
        double* d_bl = nullptr;
        if (baseline) {
            d_bl = h->d_aux + (size_t)b * chunk;
            rc = io.h2d(d_bl, baseline + e0, (size_t)ne * sizeof(double), (size_t)ne * sizeof(double), 1, bl_pinned);
            if (rc) return rc;
        }
        rc = io.inputs_ready();
        if (rc) return rc;
        const void* d_wf = h->d_in[b];
        if (encin) {
            rc = encoded_decode(h, io, *encin, 0, e0, ne);
            if (rc) return rc;
            d_wf = h->d_dec[0];
        }
        double* d_rows = h->d_rows + (size_t)b * chunk * LGDSP_NCOL;
        rc = icpc_dispatch(h, h->icpc, d_wf, sample_bytes, ne, n, d_bl, 1, 1.0, d_rows);
        if (rc) return rc;
        CK(cudaEventRecord(h->ev_free[b], h->stream));
        rc = io.d2h(out_rows + e0 * LGDSP_NCOL, d_rows, (size_t)ne * LGDSP_NCOL * sizeof(double), out_pinned);
        if (rc) return rc;
        rc = io.end_chunk();
        if (rc) return rc;
    }
    rc = io.finish();
    if (rc) return rc;
    return encin ? decode_err_check(h) : LGDSP_OK;
}

int lgdsp_icpc_run(lgdsp_handle* h, const lgdsp_icpc_params* p, const uint16_t* wf, int64_t n_events,
                   int64_t ld_samples, double* out_rows)
{
    return icpc_run_host_impl(h, p, wf, 2, nullptr, n_events, ld_samples, out_rows);
}

int lgdsp_icpc_run_ext(lgdsp_handle* h, const lgdsp_icpc_params* p, const void* wf, int32_t sample_bytes,
                       const double* baseline, int64_t n_events, int64_t ld_samples, double* out_rows)
{
    return icpc_run_host_impl(h, p, wf, sample_bytes, baseline, n_events, ld_samples, out_rows);
}

// ---- decode_data (LegendDataTypes.jl codecs; /root/reference/src/dsp_icpc.jl:313-314) ----
int64_t lgdsp_codec_max_encoded_bytes(int32_t codec, int32_t n_samples, int32_t sample_bytes)
{
    if ((codec != LGDSP_CODEC_RADWARE && codec != LGDSP_CODEC_ULEB128ZZD) || n_samples < 0) return -1;
    return codec_max_encoded_bytes(codec, n_samples, sample_bytes);
}

int lgdsp_codec_encode_host(int32_t codec, const void* wf, int32_t sample_bytes, int64_t n_events, int32_t n_samples, int64_t ld_samples,
                            int32_t shift, uint8_t* enc, int64_t enc_capacity, int64_t* offsets)
{
    if ((codec != LGDSP_CODEC_RADWARE && codec != LGDSP_CODEC_ULEB128ZZD) || !offsets || n_events < 0 || n_samples < 0 ||
        (n_events > 0 && (!wf || !enc)) || ld_samples < n_samples || (sample_bytes != 2 && sample_bytes != 4) ||
        (codec == LGDSP_CODEC_RADWARE && (sample_bytes != 2 || n_samples > 65535)))
        return LGDSP_ERR_INVALID_ARG;
    static_assert(sizeof(long long) == sizeof(int64_t), "offset type");
    const int rc = codec_encode_host(codec, wf, sample_bytes, n_events, n_samples, ld_samples, shift, enc, enc_capacity,
                                     reinterpret_cast<long long*>(offsets));
    return rc == 0 ? LGDSP_OK : (rc == -1 ? LGDSP_ERR_OOM : LGDSP_ERR_INVALID_ARG);
}

int lgdsp_decode_data_device(lgdsp_handle* h, int32_t codec, const uint8_t* d_enc, const int64_t* d_offsets, int64_t n_events,
                             int32_t n_samples, int32_t shift, void* d_wf, int32_t sample_bytes, int64_t ld_samples, int32_t* d_status)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    if (codec != LGDSP_CODEC_RADWARE && codec != LGDSP_CODEC_ULEB128ZZD) return fail(h, LGDSP_ERR_UNSUPPORTED, "unknown codec %d", codec);
    if ((codec == LGDSP_CODEC_RADWARE && sample_bytes != 2) || (sample_bytes != 2 && sample_bytes != 4))
        return fail(h, LGDSP_ERR_UNSUPPORTED, "codec %d with %d-byte samples", codec, sample_bytes);
    if (n_events < 0 || n_samples < 1 || n_samples > LGDSP_MAX_SAMPLES || ld_samples < n_samples) return fail(h, LGDSP_ERR_INVALID_ARG, "bad sizes");
    if (n_events == 0) return LGDSP_OK;
    if (!d_enc || !d_offsets || !d_wf) return fail(h, LGDSP_ERR_INVALID_ARG, "NULL pointer");
    CK(cudaEventRecord(h->ev0, h->stream));
    cudaError_t e = codec_decode_launch(codec, d_enc, reinterpret_cast<const long long*>(d_offsets), 0, n_events, n_samples, shift, d_wf,
                                        sample_bytes, ld_samples, d_status, nullptr, 0, 0 /* offsets live on the device */, h->sm_count,
                                        h->stream);
    if (e != cudaSuccess) return fail(h, LGDSP_ERR_CUDA, "decode kernel: %s", cudaGetErrorString(e));
    CK(cudaEventRecord(h->ev1, h->stream));
    h->timed = true;
    h->launches += 1;
    return LGDSP_OK;
}

int lgdsp_decode_data(lgdsp_handle* h, int32_t codec, const uint8_t* enc, const int64_t* offsets, int64_t n_events, int32_t n_samples,
                      int32_t shift, void* wf, int32_t sample_bytes, int64_t ld_samples)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    int rc = check_encoded(h, codec, enc, offsets, n_events, sample_bytes);
    if (rc) return rc;
    if (n_events < 0 || n_samples < 1 || n_samples > LGDSP_MAX_SAMPLES || ld_samples < n_samples) return fail(h, LGDSP_ERR_INVALID_ARG, "bad sizes");
    if (n_events == 0) return LGDSP_OK;
    if (!wf) return fail(h, LGDSP_ERR_INVALID_ARG, "output pointer is NULL");
    EncodedInput in{codec, sample_bytes, n_samples, shift, enc, offsets, is_pinned(enc), is_pinned(offsets)};
    const int64_t chunk = n_events < 8192 ? n_events : 8192;
    const bool out_pinned = is_pinned(wf);
    const size_t row = (size_t)n_samples * sample_bytes;
    HostIO io(h);
    rc = io.reserve((in.enc_pinned && in.off_pinned) ? 0 : encoded_stage_bytes(in, chunk, n_events), out_pinned ? 0 : (size_t)chunk * row + 256);
    if (rc) return rc;
    rc = encoded_reserve(h, in, 0, chunk, n_events);
    if (rc) return rc;
    rc = decode_err_reset(h);
    if (rc) return rc;
    int c = 0;
    for (int64_t e0 = 0; e0 < n_events; e0 += chunk, ++c) {
        const int64_t ne = (n_events - e0) < chunk ? (n_events - e0) : chunk;
        rc = io.begin_chunk(c);
        if (rc) return rc;
        if (c >= 2) CK(cudaStreamWaitEvent(h->s_copy, h->ev_free[io.b], 0));
        rc = encoded_upload(h, io, in, 0, e0, ne);
        if (rc) return rc;
        rc = io.inputs_ready();
        if (rc) return rc;
        rc = encoded_decode(h, io, in, 0, e0, ne);
        if (rc) return rc;
        CK(cudaEventRecord(h->ev_free[io.b], h->stream));
        // (d_dec is single-buffered: the D2H below is in stream order in front of the next chunk's decode kernel)
        if (ld_samples == n_samples) {
            rc = io.d2h(static_cast<char*>(wf) + (size_t)e0 * row, h->d_dec[0], (size_t)ne * row, out_pinned);
        } else {
            for (int64_t e = 0; e < ne && !rc; ++e)
                rc = io.d2h(static_cast<char*>(wf) + (size_t)(e0 + e) * ld_samples * sample_bytes, static_cast<char*>(h->d_dec[0]) + (size_t)e * row,
                            row, out_pinned);
        }
        if (rc) return rc;
        rc = io.end_chunk();
        if (rc) return rc;
    }
    rc = io.finish();
    if (rc) return rc;
    return decode_err_check(h);
}

// dsp_icpc on ENCODED waveforms: the host link carries the codec's bytes, decode_data runs on the device in front of the chain
int lgdsp_icpc_run_encoded(lgdsp_handle* h, const lgdsp_icpc_params* p, int32_t codec, const uint8_t* enc, const int64_t* offsets,
                           int32_t shift, int32_t sample_bytes, const double* baseline, int64_t n_events, double* out_rows)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    if (p) { int rc = icpc_prepare(h, p); if (rc) return rc; }
    if (!h->have_icpc) return fail(h, LGDSP_ERR_INVALID_ARG, "no parameters set");
    EncodedInput in{codec, sample_bytes, h->icpc.n, shift, enc, offsets, enc ? is_pinned(enc) : false, offsets ? is_pinned(offsets) : false};
    return icpc_run_host_impl(h, nullptr, nullptr, sample_bytes, baseline, n_events, h->icpc.n, out_rows, &in);
}

// signalstats on n_windows windows of every waveform (optionally shifted by -shift[e]): out double[n_events][n_windows][5]
int lgdsp_window_stats_run_device(lgdsp_handle* h, const void* d_wf, int32_t sample_bytes, int64_t n_events, int32_t n_samples,
                                  int64_t ld_samples, double t_first_ns, double dt_ns, const double* d_shift,
                                  const int32_t* windows, int32_t n_windows, double* d_out)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    if (sample_bytes != 2 && sample_bytes != 4) return fail(h, LGDSP_ERR_UNSUPPORTED, "sample_bytes = %d", sample_bytes);
    if (n_events < 0 || n_windows < 0 || n_windows > LGDSP_MAX_STAT_WINDOWS) return fail(h, LGDSP_ERR_INVALID_ARG, "bad n_events / n_windows");
    if (n_events == 0 || n_windows == 0) return LGDSP_OK;
    if (!d_wf || !windows || !d_out) return fail(h, LGDSP_ERR_INVALID_ARG, "NULL pointer");
    if (n_samples < 1 || ld_samples < n_samples) return fail(h, LGDSP_ERR_INVALID_ARG, "bad n_samples / ld_samples");
    if (!(dt_ns > 0) || !std::isfinite(t_first_ns)) return fail(h, LGDSP_ERR_INVALID_ARG, "bad time axis");
    for (int w = 0; w < n_windows; ++w) {
        const int a = windows[2 * w], b = windows[2 * w + 1];
        // @assert firstindex(X) <= first(idxs) <= last(idxs) <= lastindex(X)   /root/reference/src/tailstats.jl:23-25
        if (!(0 <= a && a <= b && b <= n_samples - 1)) return fail(h, LGDSP_ERR_INVALID_ARG, "window %d (%d:%d) outside the waveform", w, a, b);
    }
    if (!h->d_win) CK(cudaMalloc(&h->d_win, sizeof(int) * 2 * LGDSP_MAX_STAT_WINDOWS));
    CK(cudaMemcpyAsync(h->d_win, windows, sizeof(int) * 2 * (size_t)n_windows, cudaMemcpyHostToDevice, h->stream));
    CK(cudaEventRecord(h->ev0, h->stream));
    window_stats_launch(d_wf, sample_bytes, n_events, ld_samples, t_first_ns, dt_ns, d_shift, 1, 0xFFFFFFFFu, h->d_win, n_windows, d_out, h->stream);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, h->stream));
    h->timed = true;
    h->launches += 1;
    // the window list is read from pageable host memory by the async copy: make sure it is consumed before returning
    CK(cudaStreamSynchronize(h->stream));
    return LGDSP_OK;
}

int lgdsp_window_stats_run(lgdsp_handle* h, const void* wf, int32_t sample_bytes, int64_t n_events, int32_t n_samples,
                           int64_t ld_samples, double t_first_ns, double dt_ns, const double* shift,
                           const int32_t* windows, int32_t n_windows, double* out)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    if (sample_bytes != 2 && sample_bytes != 4) return fail(h, LGDSP_ERR_UNSUPPORTED, "sample_bytes = %d", sample_bytes);
    if (n_events < 0 || n_windows < 0 || n_windows > LGDSP_MAX_STAT_WINDOWS) return fail(h, LGDSP_ERR_INVALID_ARG, "bad n_events / n_windows");
    if (n_events == 0 || n_windows == 0) return LGDSP_OK;
    if (!wf || !out) return fail(h, LGDSP_ERR_INVALID_ARG, "NULL pointer");
    if (n_samples < 1 || ld_samples < n_samples) return fail(h, LGDSP_ERR_INVALID_ARG, "bad n_samples / ld_samples");
    const size_t sb = (size_t)sample_bytes;
    const int64_t chunk = n_events < 8192 ? n_events : 8192;
    const size_t out_per = (size_t)n_windows * 5;
    int rc = ensure_staging(h, (size_t)chunk * n_samples * sb, (size_t)chunk * out_per * sizeof(double));
    if (rc) return rc;
    if (shift) { rc = ensure_aux(h, (size_t)chunk * sizeof(double)); if (rc) return rc; }
    const unsigned char* src = static_cast<const unsigned char*>(wf);
    for (int64_t e0 = 0; e0 < n_events; e0 += chunk) {
        const int64_t ne = (n_events - e0) < chunk ? (n_events - e0) : chunk;
        CK(cudaMemcpy2DAsync(h->d_in[0], (size_t)n_samples * sb, src + (size_t)e0 * ld_samples * sb, (size_t)ld_samples * sb,
                             (size_t)n_samples * sb, (size_t)ne, cudaMemcpyHostToDevice, h->stream));
        if (shift) CK(cudaMemcpyAsync(h->d_aux, shift + e0, (size_t)ne * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        rc = lgdsp_window_stats_run_device(h, h->d_in[0], sample_bytes, ne, n_samples, n_samples, t_first_ns, dt_ns,
                                           shift ? h->d_aux : nullptr, windows, n_windows, h->d_rows);
        if (rc) return rc;
        CK(cudaMemcpyAsync(out + (size_t)e0 * out_per, h->d_rows, (size_t)ne * out_per * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    return LGDSP_OK;
}

// ---- dsp_icpc_compressed: presummed + windowed waveform per event (/root/reference/src/dsp_icpc.jl:293-499) ----
static int compressed_prepare(lgdsp_handle* h, const lgdsp_icpc_params* p_pre, const lgdsp_icpc_params* p_wdw, int pre_bytes,
                              int wdw_bytes, double presum_rate, const int32_t* aux)
{
    if (p_pre) { int rc = icpc_prepare(h, p_pre, 0); if (rc) return rc; }
    if (p_wdw) { int rc = icpc_prepare(h, p_wdw, 1); if (rc) return rc; }
    if (!h->have_icpc || !h->have_icpc_w) return fail(h, LGDSP_ERR_INVALID_ARG, "dsp_icpc_compressed: parameters of both waveforms are needed");
    if (!(presum_rate >= 1.0)) return fail(h, LGDSP_ERR_INVALID_ARG, "presum_rate must be >= 1");
    if ((pre_bytes != 2 && pre_bytes != 4) || (wdw_bytes != 2 && wdw_bytes != 4))
        return fail(h, LGDSP_ERR_UNSUPPORTED, "sample_bytes: uint16 (2) or uint32 (4) samples");
    if (!aux) return fail(h, LGDSP_ERR_INVALID_ARG, "aux_windows is NULL");
    const int n = h->icpc.n;
    for (int w = 0; w < 4; ++w) {
        const int a = aux[2 * w], b = aux[2 * w + 1];
        if (!(0 <= a && a <= b && b <= n - 1)) return fail(h, LGDSP_ERR_INVALID_ARG, "auxiliary window %d (%d:%d) outside the presummed waveform", w, a, b);
    }
    if (!h->d_win) CK(cudaMalloc(&h->d_win, sizeof(int) * 2 * LGDSP_MAX_STAT_WINDOWS));
    // window order of the statistics output: auxbl1, auxbl2, bl_window (raw), auxpz1, auxpz2 (baseline-subtracted)
    const int win[10] = {aux[0], aux[1], aux[2], aux[3], h->icpc.bl_from, h->icpc.bl_until, aux[4], aux[5], aux[6], aux[7]};
    CK(cudaMemcpyAsync(h->d_win, win, sizeof(win), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return LGDSP_OK;
}

// the three launches of one batch, in stream order
static int compressed_launch(lgdsp_handle* h, const void* d_pre, int pre_bytes, int64_t ld_pre, const void* d_wdw, int wdw_bytes,
                              int64_t ld_wdw, double presum_rate, int64_t ne, double* d_rows_pre, double* d_rows_wdw, double* d_stats)
{
    // :332-346, :356-375, :383, :397-428, :439-455 on the presummed waveform
    int rc = icpc_dispatch(h, h->icpc, d_pre, pre_bytes, ne, ld_pre, nullptr, 1, 1.0, d_rows_pre);
    if (rc) return rc;
    // :338-339, :346 (slope_residual_sigma), :365-366: windows 3, 4 are shifted by the baseline mean of the first launch
    window_stats_launch(d_pre, pre_bytes, ne, ld_pre, h->icpc.t_first, h->icpc.dt, d_rows_pre + LGDSP_COL_blmean, LGDSP_NCOL, 0x18u,
                        h->d_win, 5, d_stats, h->stream);
    // :350 the windowed waveform is shifted by -blmean / presum_rate; :378-394, :431-435, :458
    rc = icpc_dispatch(h, h->icpc_w, d_wdw, wdw_bytes, ne, ld_wdw, d_rows_pre + LGDSP_COL_blmean, LGDSP_NCOL, presum_rate, d_rows_wdw);
    if (rc) return rc;
    h->launches += 1;
    return LGDSP_OK;
}

int lgdsp_icpc_compressed_run_device(lgdsp_handle* h, const lgdsp_icpc_params* p_pre, const lgdsp_icpc_params* p_wdw,
                                     const void* d_wf_pre, int32_t pre_sample_bytes, int64_t ld_pre, const void* d_wf_wdw,
                                     int32_t wdw_sample_bytes, int64_t ld_wdw, double presum_rate, const int32_t* aux_windows,
                                     int64_t n_events, double* d_rows_pre, double* d_rows_wdw, double* d_stats)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    int rc = compressed_prepare(h, p_pre, p_wdw, pre_sample_bytes, wdw_sample_bytes, presum_rate, aux_windows);
    if (rc) return rc;
    rc = check_wf(h, d_wf_pre, n_events, ld_pre, h->icpc.n, true, pre_sample_bytes);
    if (rc) return rc;
    rc = check_wf(h, d_wf_wdw, n_events, ld_wdw, h->icpc_w.n, true, wdw_sample_bytes);
    if (rc) return rc;
    if (n_events == 0) return LGDSP_OK;
    if (!d_rows_pre || !d_rows_wdw || !d_stats) return fail(h, LGDSP_ERR_INVALID_ARG, "output pointer is NULL");
    CK(cudaEventRecord(h->ev0, h->stream));
    rc = compressed_launch(h, d_wf_pre, pre_sample_bytes, ld_pre, d_wf_wdw, wdw_sample_bytes, ld_wdw, presum_rate, n_events, d_rows_pre,
                           d_rows_wdw, d_stats);
    if (rc) return rc;
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, h->stream));
    h->timed = true;
    return LGDSP_OK;
}

// host buffers: raw samples (wf_* != NULL) or encoded waveform sets (enc_* != NULL) per waveform kind
static int compressed_run_host_impl(lgdsp_handle* h, const lgdsp_icpc_params* p_pre, const lgdsp_icpc_params* p_wdw, const void* wf_pre,
                                    int pre_sample_bytes, int64_t ld_pre, const void* wf_wdw, int wdw_sample_bytes, int64_t ld_wdw,
                                    const EncodedInput* enc_pre, const EncodedInput* enc_wdw, double presum_rate,
                                    const int32_t* aux_windows, int64_t n_events, double* rows_pre, double* rows_wdw, double* stats)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    int rc = compressed_prepare(h, p_pre, p_wdw, pre_sample_bytes, wdw_sample_bytes, presum_rate, aux_windows);
    if (rc) return rc;
    const int np = h->icpc.n, nw = h->icpc_w.n;
    EncodedInput ep{}, ew{};
    if (enc_pre) { ep = *enc_pre; ep.n_samples = np; rc = check_encoded(h, ep.codec, ep.enc, ep.offsets, n_events, pre_sample_bytes); }
    else rc = check_wf(h, wf_pre, n_events, ld_pre, np, false, pre_sample_bytes);
    if (rc) return rc;
    if (enc_wdw) { ew = *enc_wdw; ew.n_samples = nw; rc = check_encoded(h, ew.codec, ew.enc, ew.offsets, n_events, wdw_sample_bytes); }
    else rc = check_wf(h, wf_wdw, n_events, ld_wdw, nw, false, wdw_sample_bytes);
    if (rc) return rc;
    if (n_events < 0) return fail(h, LGDSP_ERR_INVALID_ARG, "n_events < 0");
    if (n_events == 0) return LGDSP_OK;
    if (!rows_pre || !rows_wdw || !stats) return fail(h, LGDSP_ERR_INVALID_ARG, "output pointer is NULL");
    const int64_t chunk = n_events < h->host_chunk ? n_events : h->host_chunk;
    const size_t sbp = (size_t)pre_sample_bytes, sbw = (size_t)wdw_sample_bytes;
    const size_t need[2] = {enc_pre ? 16 : (size_t)chunk * np * sbp, enc_wdw ? 16 : (size_t)chunk * nw * sbw};
    for (int k = 0; k < 2; ++k)
        if (need[k] > h->cin_cap[k]) {
            CK(cudaStreamSynchronize(h->stream));
            CK(cudaStreamSynchronize(h->s_copy));
            for (int b = 0; b < 2; ++b) { cudaFree(h->d_cin[b][k]); h->d_cin[b][k] = nullptr; }
            h->cin_cap[k] = 0;
            for (int b = 0; b < 2; ++b) CK(cudaMalloc(&h->d_cin[b][k], need[k]));
            h->cin_cap[k] = need[k];
        }
    const size_t per_event = 2 * LGDSP_NCOL + 5 * LGDSP_NSTAT;
    if ((size_t)2 * chunk * per_event * sizeof(double) > h->crows_cap) {
        CK(cudaStreamSynchronize(h->stream));
        cudaFree(h->d_crows); h->d_crows = nullptr; h->crows_cap = 0;
        CK(cudaMalloc(&h->d_crows, (size_t)2 * chunk * per_event * sizeof(double)));
        h->crows_cap = (size_t)2 * chunk * per_event * sizeof(double);
    }
    const bool pin_pre = enc_pre ? (ep.enc_pinned && ep.off_pinned) : is_pinned(wf_pre);
    const bool pin_wdw = enc_wdw ? (ew.enc_pinned && ew.off_pinned) : is_pinned(wf_wdw);
    const bool pin_out = is_pinned(rows_pre) && is_pinned(rows_wdw) && is_pinned(stats);
    HostIO io(h);
    size_t stage_in = 0;
    if (!pin_pre) stage_in += (enc_pre ? encoded_stage_bytes(ep, chunk, n_events) : (size_t)chunk * np * sbp) + 512;
    if (!pin_wdw) stage_in += (enc_wdw ? encoded_stage_bytes(ew, chunk, n_events) : (size_t)chunk * nw * sbw) + 512;
    rc = io.reserve(stage_in, pin_out ? 0 : (size_t)chunk * per_event * sizeof(double) + 1024);
    if (rc) return rc;
    if (enc_pre) { rc = encoded_reserve(h, ep, 0, chunk, n_events); if (rc) return rc; }
    if (enc_wdw) { rc = encoded_reserve(h, ew, 1, chunk, n_events); if (rc) return rc; }
    if (enc_pre || enc_wdw) { rc = decode_err_reset(h); if (rc) return rc; }
    const unsigned char* sp = static_cast<const unsigned char*>(wf_pre);
    const unsigned char* sw = static_cast<const unsigned char*>(wf_wdw);
    int c = 0;
    for (int64_t e0 = 0; e0 < n_events; e0 += chunk, ++c) {
        const int64_t ne = (n_events - e0) < chunk ? (n_events - e0) : chunk;
        rc = io.begin_chunk(c);
        if (rc) return rc;
        const int b = io.b;
        if (c >= 2) CK(cudaStreamWaitEvent(h->s_copy, h->ev_free[b], 0));
        if (enc_pre) rc = encoded_upload(h, io, ep, 0, e0, ne);
        else rc = io.h2d(h->d_cin[b][0], sp + (size_t)e0 * ld_pre * sbp, (size_t)ld_pre * sbp, (size_t)np * sbp, (size_t)ne, pin_pre);
        if (rc) return rc;
        if (enc_wdw) rc = encoded_upload(h, io, ew, 1, e0, ne);
        else rc = io.h2d(h->d_cin[b][1], sw + (size_t)e0 * ld_wdw * sbw, (size_t)ld_wdw * sbw, (size_t)nw * sbw, (size_t)ne, pin_wdw);
        if (rc) return rc;
        rc = io.inputs_ready();
        if (rc) return rc;
        const void* d_pre = h->d_cin[b][0];
        const void* d_wdw = h->d_cin[b][1];
        if (enc_pre) { rc = encoded_decode(h, io, ep, 0, e0, ne); if (rc) return rc; d_pre = h->d_dec[0]; }
        if (enc_wdw) { rc = encoded_decode(h, io, ew, 1, e0, ne); if (rc) return rc; d_wdw = h->d_dec[1]; }
        double* d_rp = h->d_crows + (size_t)b * chunk * per_event;
        double* d_rw = d_rp + (size_t)chunk * LGDSP_NCOL;
        double* d_st = d_rw + (size_t)chunk * LGDSP_NCOL;
        rc = compressed_launch(h, d_pre, pre_sample_bytes, np, d_wdw, wdw_sample_bytes, nw, presum_rate, ne, d_rp, d_rw, d_st);
        if (rc) return rc;
        CK(cudaGetLastError());
        CK(cudaEventRecord(h->ev_free[b], h->stream));
        rc = io.d2h(rows_pre + e0 * LGDSP_NCOL, d_rp, (size_t)ne * LGDSP_NCOL * sizeof(double), pin_out);
        if (!rc) rc = io.d2h(rows_wdw + e0 * LGDSP_NCOL, d_rw, (size_t)ne * LGDSP_NCOL * sizeof(double), pin_out);
        if (!rc) rc = io.d2h(stats + e0 * 5 * LGDSP_NSTAT, d_st, (size_t)ne * 5 * LGDSP_NSTAT * sizeof(double), pin_out);
        if (rc) return rc;
        rc = io.end_chunk();
        if (rc) return rc;
    }
    rc = io.finish();
    if (rc) return rc;
    return (enc_pre || enc_wdw) ? decode_err_check(h) : LGDSP_OK;
}

int lgdsp_icpc_compressed_run(lgdsp_handle* h, const lgdsp_icpc_params* p_pre, const lgdsp_icpc_params* p_wdw, const void* wf_pre,
                              int32_t pre_sample_bytes, int64_t ld_pre, const void* wf_wdw, int32_t wdw_sample_bytes,
                              int64_t ld_wdw, double presum_rate, const int32_t* aux_windows, int64_t n_events, double* rows_pre,
                              double* rows_wdw, double* stats)
{
    return compressed_run_host_impl(h, p_pre, p_wdw, wf_pre, pre_sample_bytes, ld_pre, wf_wdw, wdw_sample_bytes, ld_wdw, nullptr, nullptr,
                                    presum_rate, aux_windows, n_events, rows_pre, rows_wdw, stats);
}

// dsp_icpc_compressed with decode_data on the device (/root/reference/src/dsp_icpc.jl:313-314): both waveform kinds arrive
// as encoded byte streams in host memory
int lgdsp_icpc_compressed_run_encoded(lgdsp_handle* h, const lgdsp_icpc_params* p_pre, const lgdsp_icpc_params* p_wdw,
                                      int32_t pre_codec, const uint8_t* enc_pre, const int64_t* off_pre, int32_t pre_shift,
                                      int32_t pre_sample_bytes, int32_t wdw_codec, const uint8_t* enc_wdw, const int64_t* off_wdw,
                                      int32_t wdw_shift, int32_t wdw_sample_bytes, double presum_rate, const int32_t* aux_windows,
                                      int64_t n_events, double* rows_pre, double* rows_wdw, double* stats)
{
    EncodedInput ep{pre_codec, pre_sample_bytes, 0, pre_shift, enc_pre, off_pre, enc_pre ? is_pinned(enc_pre) : false,
                    off_pre ? is_pinned(off_pre) : false};
    EncodedInput ew{wdw_codec, wdw_sample_bytes, 0, wdw_shift, enc_wdw, off_wdw, enc_wdw ? is_pinned(enc_wdw) : false,
                    off_wdw ? is_pinned(off_wdw) : false};
    return compressed_run_host_impl(h, p_pre, p_wdw, nullptr, pre_sample_bytes, 0, nullptr, wdw_sample_bytes, 0, &ep, &ew, presum_rate,
                                    aux_windows, n_events, rows_pre, rows_wdw, stats);
}

// debug: per-phase cycle counters of the last lgdsp_icpc_run_device call (sum over CTAs; index 0: TMA wait,
// 1..6: P1, P2, P3, P4a, P4b, P5, see lgdsp_icpc.cu).  enable != 0 switches the counters on for later calls.
int lgdsp_debug_phase_cycles(lgdsp_handle* h, int enable, double* out8)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    if (out8) {
        for (int i = 0; i < 8; ++i) out8[i] = 0.0;
        if (h->d_phase && h->phase_grid > 0) {
            CK(cudaStreamSynchronize(h->stream));
            std::vector<unsigned long long> v((size_t)h->phase_grid * 8);
            CK(cudaMemcpy(v.data(), h->d_phase, v.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            for (int b = 0; b < h->phase_grid; ++b)
                for (int i = 0; i < 8; ++i) out8[i] += (double)v[(size_t)b * 8 + i];
        }
    }
    h->phase_on = enable != 0;
    return LGDSP_OK;
}

// debug (library built with -DLGDSP_PROFILE_SECTIONS only): cycles per (section, warp) of the last run, out[32*8]
int lgdsp_debug_section_cycles(lgdsp_handle* h, double* out256)
{
    if (!h || !out256) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    for (int i = 0; i < 256; ++i) out256[i] = 0.0;
    if (!h->d_phase || h->phase_grid <= 0) return LGDSP_OK;
    CK(cudaStreamSynchronize(h->stream));
    unsigned long long v[256];
    CK(cudaMemcpy(v, h->d_phase + (size_t)h->phase_grid * 8, sizeof(v), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 256; ++i) out256[i] = (double)v[i];
    return LGDSP_OK;
}

// ---------------------------------------------------------------------------------------------------
// trapezoid sweeps
// ---------------------------------------------------------------------------------------------------
// variants -> device table (filter taps are differenced onto the prefix sums TT and uploaded)
static int sweep_prepare(lgdsp_handle* h, const lgdsp_sweep_params* p, const lgdsp_sweep_variant* variants, int32_t nvar,
                         SweepDev& D)
{
    if (!p || !variants) return fail(h, LGDSP_ERR_INVALID_ARG, "sweep params/variants NULL");
    if (p->struct_size != sizeof(lgdsp_sweep_params) || p->version != LGDSP_PARAMS_VERSION)
        return fail(h, LGDSP_ERR_INVALID_ARG, "lgdsp_sweep_params: size/version mismatch");
    const int n = p->n_samples;
    if (n < 64 || n > LGDSP_MAX_SAMPLES || n % 8 != 0) return fail(h, LGDSP_ERR_UNSUPPORTED, "n_samples = %d unsupported", n);
    if (nvar < 1 || nvar > 1024) return fail(h, LGDSP_ERR_UNSUPPORTED, "n_variants = %d: need 1..1024", nvar);
    if (!(0 <= p->bl_from && p->bl_from <= p->bl_until && p->bl_until < n)) return fail(h, LGDSP_ERR_INVALID_ARG, "bl_window outside the waveform");
    const lgdsp_dni& d = p->sig_dni;
    if (!(d.degree >= 0 && d.degree <= LGDSP_MAX_DNI_DEG && d.n_w > d.degree && d.n_w <= LGDSP_MAX_DNI))
        return fail(h, LGDSP_ERR_UNSUPPORTED, "PolynomialDNI outside the supported range");
    if (p->tx_min_n < 1) return fail(h, LGDSP_ERR_INVALID_ARG, "tx_min_n < 1");
    std::vector<SweepVar> sv(nvar);
    std::vector<double> taps;            // all differenced tap arrays, back to back
    std::vector<size_t> tap_off(nvar, 0);
    for (int v = 0; v < nvar; ++v) {
        const lgdsp_sweep_variant& in = variants[v];
        SweepVar& o = sv[v];
        memset(&o, 0, sizeof(o));
        o.kind = in.kind;
        o.pick_ns = in.pickoff_ns;
        o.mode = in.pickoff_mode;
        if (in.kind == 0) {
            if (!make_trap(in.trap, n, d.n_w, o.t))
                return fail(h, LGDSP_ERR_INVALID_ARG, "variant %d: trapezoid (%d,%d,%d) leaves fewer outputs than the DNI window", v,
                            in.trap.navg, in.trap.ngap, in.trap.navg2);
            o.L = o.t.L;
        } else if (in.kind == 1) {
            const int L = in.n_taps;
            if (!in.coeffs || L < 2 || L > LGDSP_MAX_FIR || n - L + 1 < d.n_w)
                return fail(h, LGDSP_ERR_INVALID_ARG, "variant %d: FIR with %d taps does not fit (need >= %d outputs)", v, L, d.n_w);
            o.L = L;
            tap_off[v] = taps.size();
            // out[j] = sum_k c[k] y[j+L-1-k] = sum_{k=0}^{L} g[k] TT[j+L-k],  g = first difference of c
            taps.push_back(in.coeffs[0]);
            for (int k = 1; k < L; ++k) taps.push_back(in.coeffs[k] - in.coeffs[k - 1]);
            taps.push_back(-in.coeffs[L - 1]);
        } else if (in.kind == 2) {
            const int T = in.n_taps;
            if (!in.coeffs || T < 1 || T > LGDSP_MAX_SG || T > n || in.sg_offset < 0 || in.sg_offset >= T)
                return fail(h, LGDSP_ERR_INVALID_ARG, "variant %d: bad Savitzky-Golay tap count/offset", v);
            if (!(0 <= in.win_from && in.win_from <= in.win_until && in.win_until <= n - T))
                return fail(h, LGDSP_ERR_INVALID_ARG, "variant %d: current_window %d:%d outside the SG trace", v, in.win_from, in.win_until);
            o.L = T; o.sg_off = in.sg_offset; o.win_from = in.win_from; o.win_until = in.win_until;
            tap_off[v] = taps.size();
            // s[j] = sum_k h[k] y[j+k] = sum_{k=0}^{T} gg[k] TT[j+k]
            taps.push_back(-in.coeffs[0]);
            for (int k = 1; k < T; ++k) taps.push_back(in.coeffs[k - 1] - in.coeffs[k]);
            taps.push_back(in.coeffs[T - 1]);
        } else {
            return fail(h, LGDSP_ERR_INVALID_ARG, "variant %d: unknown kind %d", v, in.kind);
        }
    }
    if (nvar > h->vars_cap) {
        cudaFree(h->d_vars); h->d_vars = nullptr; h->vars_cap = 0;
        CK(cudaMalloc(&h->d_vars, sizeof(SweepVar) * nvar));
        h->vars_cap = nvar;
    }
    if (taps.size() > h->taps_cap) {
        cudaFree(h->d_taps); h->d_taps = nullptr; h->taps_cap = 0;
        CK(cudaMalloc(&h->d_taps, sizeof(double) * taps.size()));
        h->taps_cap = taps.size();
    }
    for (int v = 0; v < nvar; ++v)
        if (sv[v].kind != 0) sv[v].g = h->d_taps + tap_off[v];
    if (!taps.empty()) CK(cudaMemcpyAsync(h->d_taps, taps.data(), sizeof(double) * taps.size(), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_vars, sv.data(), sizeof(SweepVar) * nvar, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_sweep_dniA, d.A, sizeof(double) * LGDSP_MAX_DNI * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    D.n = n; D.tx_min_n = p->tx_min_n; D.t_first = p->t_first_ns; D.dt = p->dt_ns;
    D.bl_from = p->bl_from; D.bl_until = p->bl_until; D.km1 = p->pz_km1;
    D.sig_dni.n_w = d.n_w; D.sig_dni.m = d.degree + 1;
    D.dni_A = h->d_sweep_dniA; D.vars = h->d_vars; D.nvar = nvar; D.out_f64 = p->out_f64 ? 1 : 0;
    D.n_other = 0;
    for (int v = 0; v < nvar; ++v) D.n_other += (sv[v].kind != 0);
    // Reach of the variant set around the crossing sample `pos` of t50 (t50 lies in [pos-1, pos]):
    // from = rint(pc) - n_w/2, pc = (t50 - t_first)/dt + pick/dt - (L-1)  (dni_window), look-ups in [from, from + L + n_w)
    D.warp_ok = 0;
    {
        int ex = 0;
        D.dt_pow2 = (p->dt_ns > 0 && std::frexp(p->dt_ns, &ex) == 0.5) ? 1 : 0;
        D.rdt = 1.0 / p->dt_ns;
    }
    if (D.n_other == 0) {
        bool all1 = true, all0 = true;
        double lo = 1e300, hi = -1e300;
        for (int v = 0; v < nvar; ++v) {
            all1 = all1 && sv[v].mode != 0;
            all0 = all0 && sv[v].mode == 0;
            const double rel = sv[v].pick_ns / p->dt_ns - (double)(sv[v].L - 1) - (double)(d.n_w / 2);
            const double base = sv[v].mode ? rel : rel - p->t_first_ns / p->dt_ns;
            lo = std::min(lo, std::floor(base - 1.5) - 2.0);
            hi = std::max(hi, std::ceil(base + 0.5) + (double)(sv[v].L + d.n_w) + 2.0);
        }
        if (all1) { lo = std::min(lo, -3.0); hi = std::max(hi, 3.0); }
        const int wmax = sweep_warp_max_window();
        int longest = 0;
        for (int v = 0; v < nvar; ++v) longest = std::max(longest, sv[v].L + d.n_w + 2);
        if ((all1 || all0) && wmax > 0 && hi - lo <= (double)wmax && longest <= wmax && std::fabs(lo) < 1e9) {
            const int need = std::max((int)(hi - lo), longest);
            D.win_steps = std::max(4, (need + 287) / 288);   // (>= 4: the window area also holds pass 1's group table)
            D.win_mode = all1 ? 1 : 0;
            D.win_rel_lo = all1 ? (int)lo : 0;
            D.win_abs_lo = all1 ? 0 : std::max(0, (int)lo);
            D.warp_ok = D.win_steps * 288 <= wmax ? 1 : 0;
        }
        // fixed pick-offs: the windows are the same for every event -- the exact last look-up of the set (dni_window's arithmetic)
        D.stream_n = n;
        if (all0) {
            long long last = p->bl_until + 1;
            for (int v = 0; v < nvar; ++v) {
                const int Lf = sv[v].L, nout = n - Lf + 1;
                const double tf = std::fma((double)(Lf - 1), p->dt_ns, p->t_first_ns);
                double pc = (sv[v].pick_ns - tf) / p->dt_ns;
                if (!(pc >= 0)) pc = 0;
                if (pc > nout - 1) pc = nout - 1;
                long long f = (long long)std::nearbyint(pc) - d.n_w / 2;
                if (f < 0) f = 0;
                if (f > nout - d.n_w) f = nout - d.n_w;
                last = std::max(last, f + Lf + d.n_w);
            }
            D.stream_n = (int)std::min<long long>(n, last + 8);
        }
    }
    D.bl_inv_n = 1.0 / (double)(p->bl_until - p->bl_from + 1);
    {
        typedef long double ld;
        const int a = p->bl_from, b = p->bl_until;
        const ld t0 = p->t_first_ns, dt = p->dt_ns, cnt = (ld)(b - a + 1);
        auto s2 = [](ld k) { return k * (k + 1.0L) * (2.0L * k + 1.0L) / 6.0L; };
        const ld si = 0.5L * (ld)(a + b) * cnt, sii = s2((ld)b) - s2((ld)a - 1.0L);
        D.bl_sX = (double)(cnt * t0 + dt * si);
        D.bl_sXX = (double)(cnt * t0 * t0 + 2.0L * t0 * dt * si + dt * dt * sii);
    }
    return LGDSP_OK;
}

static int sweep_run_device_impl(lgdsp_handle* h, const lgdsp_sweep_params* p, const void* d_wf, int sample_bytes,
                                 const double* d_baseline, int64_t n_events, int64_t ld_samples,
                                 const lgdsp_sweep_variant* variants, int32_t n_variants, void* d_out, double* d_aux)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    SweepDev D{};
    int rc = sweep_prepare(h, p, variants, n_variants, D);
    if (rc) return rc;
    rc = check_wf(h, d_wf, n_events, ld_samples, D.n, true, sample_bytes);
    if (rc) return rc;
    if (n_events == 0) return LGDSP_OK;
    if (!d_out) return fail(h, LGDSP_ERR_INVALID_ARG, "output pointer is NULL");
    const long long cap = (long long)h->sm_count * h->sweep_bps;
    const int grid = (int)(n_events < cap ? n_events : cap);
    CK(cudaEventRecord(h->ev0, h->stream));
    if (sweep_launch(D, p->sig_dni.A, d_wf, sample_bytes, n_events, ld_samples, d_baseline, d_out, d_aux, grid, h->sm_count, h->stream) < 0)
        return fail(h, LGDSP_ERR_UNSUPPORTED, "LGDSP_SWEEP_PATH=warp: this variant set / sample type runs on the one-CTA-per-waveform kernel");
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, h->stream));
    h->timed = true;
    h->launches += 1;
    return LGDSP_OK;
}

static int sweep_run_host_impl(lgdsp_handle* h, const lgdsp_sweep_params* p, const void* wf, int sample_bytes,
                               const double* baseline, int64_t n_events, int64_t ld_samples,
                               const lgdsp_sweep_variant* variants, int32_t n_variants, void* out, double* aux)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    SweepDev D{};
    int rc = sweep_prepare(h, p, variants, n_variants, D);
    if (rc) return rc;
    const int n = D.n;
    rc = check_wf(h, wf, n_events, ld_samples, n, false, sample_bytes);
    if (rc) return rc;
    if (n_events == 0) return LGDSP_OK;
    if (!out) return fail(h, LGDSP_ERR_INVALID_ARG, "output pointer is NULL");
    const size_t sb = (size_t)sample_bytes;
    const int64_t chunk = n_events < 8192 ? n_events : 8192;
    rc = ensure_staging(h, (size_t)chunk * n * sb, 0);
    if (rc) return rc;
    if (baseline) {
        rc = ensure_aux(h, (size_t)2 * chunk * sizeof(double));
        if (rc) return rc;
    }
    const size_t esz = D.out_f64 ? sizeof(double) : sizeof(float);
    const size_t ob = (size_t)2 * chunk * (n_variants * esz + 4 * sizeof(double));
    if (ob > h->sweep_out_cap) {
        cudaFree(h->d_sweep_out); h->d_sweep_out = nullptr; h->sweep_out_cap = 0;
        CK(cudaMalloc(&h->d_sweep_out, ob));
        h->sweep_out_cap = ob;
    }
    const long long cap = (long long)h->sm_count * h->sweep_bps;
    char* base = reinterpret_cast<char*>(h->d_sweep_out);
    const unsigned char* src = static_cast<const unsigned char*>(wf);
    const size_t out_bytes = (size_t)chunk * n_variants * esz;   // per buffer; the aux buffers follow the two output buffers
    int c = 0;
    for (int64_t e0 = 0; e0 < n_events; e0 += chunk, ++c) {
        const int b = c & 1;
        const int64_t ne = (n_events - e0) < chunk ? (n_events - e0) : chunk;
        if (c >= 2) CK(cudaStreamWaitEvent(h->s_copy, h->ev_free[b], 0));
        CK(cudaMemcpy2DAsync(h->d_in[b], (size_t)n * sb, src + (size_t)e0 * ld_samples * sb, (size_t)ld_samples * sb, (size_t)n * sb,
                             (size_t)ne, cudaMemcpyHostToDevice, h->s_copy));
        double* d_bl = nullptr;
        if (baseline) {
            d_bl = h->d_aux + (size_t)b * chunk;
            CK(cudaMemcpyAsync(d_bl, baseline + e0, (size_t)ne * sizeof(double), cudaMemcpyHostToDevice, h->s_copy));
        }
        CK(cudaEventRecord(h->ev_ready[b], h->s_copy));
        CK(cudaStreamWaitEvent(h->stream, h->ev_ready[b], 0));
        void* d_o = base + (size_t)b * out_bytes;
        double* d_a = aux ? reinterpret_cast<double*>(base + 2 * out_bytes) + (size_t)b * chunk * 4 : nullptr;
        const int grid = (int)(ne < cap ? ne : cap);
        if (sweep_launch(D, p->sig_dni.A, h->d_in[b], sample_bytes, ne, n, d_bl, d_o, d_a, grid, h->sm_count, h->stream) < 0)
            return fail(h, LGDSP_ERR_UNSUPPORTED, "LGDSP_SWEEP_PATH=warp: this variant set / sample type runs on the one-CTA-per-waveform kernel");
        CK(cudaGetLastError());
        h->launches += 1;
        CK(cudaEventRecord(h->ev_free[b], h->stream));
        CK(cudaMemcpyAsync(reinterpret_cast<char*>(out) + (size_t)e0 * n_variants * esz, d_o, (size_t)ne * n_variants * esz,
                           cudaMemcpyDeviceToHost, h->stream));
        if (aux) CK(cudaMemcpyAsync(aux + e0 * 4, d_a, (size_t)ne * 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->s_copy));
    CK(cudaStreamSynchronize(h->stream));
    return LGDSP_OK;
}

int lgdsp_sweep_run_device(lgdsp_handle* h, const lgdsp_sweep_params* p, const uint16_t* d_wf, int64_t n_events,
                           int64_t ld_samples, const lgdsp_sweep_variant* variants, int32_t n_variants, void* d_out,
                           double* d_aux)
{
    return sweep_run_device_impl(h, p, d_wf, 2, nullptr, n_events, ld_samples, variants, n_variants, d_out, d_aux);
}

int lgdsp_sweep_run(lgdsp_handle* h, const lgdsp_sweep_params* p, const uint16_t* wf, int64_t n_events, int64_t ld_samples,
                    const lgdsp_sweep_variant* variants, int32_t n_variants, void* out, double* aux)
{
    return sweep_run_host_impl(h, p, wf, 2, nullptr, n_events, ld_samples, variants, n_variants, out, aux);
}

int lgdsp_sweep_run_ext_device(lgdsp_handle* h, const lgdsp_sweep_params* p, const void* d_wf, int32_t sample_bytes,
                               const double* d_baseline, int64_t n_events, int64_t ld_samples,
                               const lgdsp_sweep_variant* variants, int32_t n_variants, void* d_out, double* d_aux)
{
    return sweep_run_device_impl(h, p, d_wf, sample_bytes, d_baseline, n_events, ld_samples, variants, n_variants, d_out, d_aux);
}

int lgdsp_sweep_run_ext(lgdsp_handle* h, const lgdsp_sweep_params* p, const void* wf, int32_t sample_bytes, const double* baseline,
                        int64_t n_events, int64_t ld_samples, const lgdsp_sweep_variant* variants, int32_t n_variants, void* out,
                        double* aux)
{
    return sweep_run_host_impl(h, p, wf, sample_bytes, baseline, n_events, ld_samples, variants, n_variants, out, aux);
}

// the trapezoid-only entry points (kept for callers of the first ABI revision): thin wrappers
static std::vector<lgdsp_sweep_variant> trap_to_general(const lgdsp_trap_variant* variants, int32_t n)
{
    std::vector<lgdsp_sweep_variant> v((size_t)(n > 0 ? n : 0));
    for (int i = 0; i < n; ++i) {
        memset(&v[i], 0, sizeof(v[i]));
        v[i].kind = 0;
        v[i].pickoff_mode = variants[i].pickoff_mode;
        v[i].pickoff_ns = variants[i].pickoff_ns;
        v[i].trap = variants[i].trap;
    }
    return v;
}

int lgdsp_trap_sweep_run_device(lgdsp_handle* h, const lgdsp_sweep_params* p, const uint16_t* d_wf, int64_t n_events,
                                int64_t ld_samples, const lgdsp_trap_variant* variants, int32_t n_variants, float* d_out)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    if (!variants) return fail(h, LGDSP_ERR_INVALID_ARG, "sweep params/variants NULL");
    if (p && p->out_f64) return fail(h, LGDSP_ERR_INVALID_ARG, "lgdsp_trap_sweep_run_device writes float: out_f64 must be 0");
    std::vector<lgdsp_sweep_variant> v = trap_to_general(variants, n_variants);
    return lgdsp_sweep_run_device(h, p, d_wf, n_events, ld_samples, v.data(), n_variants, d_out, nullptr);
}

int lgdsp_trap_sweep_run(lgdsp_handle* h, const lgdsp_sweep_params* p, const uint16_t* wf, int64_t n_events,
                         int64_t ld_samples, const lgdsp_trap_variant* variants, int32_t n_variants, float* out)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    if (!variants) return fail(h, LGDSP_ERR_INVALID_ARG, "sweep params/variants NULL");
    if (p && p->out_f64) return fail(h, LGDSP_ERR_INVALID_ARG, "lgdsp_trap_sweep_run writes float: out_f64 must be 0");
    std::vector<lgdsp_sweep_variant> v = trap_to_general(variants, n_variants);
    return lgdsp_sweep_run(h, p, wf, n_events, ld_samples, v.data(), n_variants, out, nullptr);
}

// ---------------------------------------------------------------------------------------------------
// dsp_sipm  (/root/reference/src/dsp_sipm.jl:47-158)
// ---------------------------------------------------------------------------------------------------
static int sipm_prepare(lgdsp_handle* h, const lgdsp_sipm_params* p, SipmDev& D, int& bps)
{
    if (!p) return fail(h, LGDSP_ERR_INVALID_ARG, "params is NULL");
    if (p->struct_size != sizeof(lgdsp_sipm_params) || p->version != LGDSP_PARAMS_VERSION)
        return fail(h, LGDSP_ERR_INVALID_ARG, "lgdsp_sipm_params: size/version mismatch (got %u/%u, want %zu/%u)", p->struct_size,
                    p->version, sizeof(lgdsp_sipm_params), LGDSP_PARAMS_VERSION);
    const int n = p->n_samples;
    if (n < 16 || n > LGDSP_MAX_SAMPLES) return fail(h, LGDSP_ERR_UNSUPPORTED, "n_samples = %d: need 16 .. %d", n, LGDSP_MAX_SAMPLES);
    if (p->sample_kind != LGDSP_SAMPLE_U16 && p->sample_kind != LGDSP_SAMPLE_F32)
        return fail(h, LGDSP_ERR_UNSUPPORTED, "sample_kind = %d: LGDSP_SAMPLE_U16 or LGDSP_SAMPLE_F32", p->sample_kind);
    if (!(p->dt_ns > 0) || !std::isfinite(p->t_first_ns)) return fail(h, LGDSP_ERR_INVALID_ARG, "bad time axis");
    if (!(0 <= p->trunc_from && p->trunc_from <= p->trunc_until && p->trunc_until <= n - 1))
        return fail(h, LGDSP_ERR_INVALID_ARG, "t0_hpge_window %d:%d outside the waveform", p->trunc_from, p->trunc_until);
    const lgdsp_sg& sg = p->sg;
    if (sg.n_taps < 1 || sg.n_taps > LGDSP_MAX_SG || sg.n_taps > n || sg.offset < 0 || sg.offset >= sg.n_taps)
        return fail(h, LGDSP_ERR_INVALID_ARG, "sg: bad tap count/offset");
    const int n_sg = n - sg.n_taps + 1;
    const lgdsp_trap& t = p->trap;
    if (t.navg < 1 || t.navg2 < 1 || t.ngap < 0 || t.navg + t.ngap + t.navg2 > n_sg)
        return fail(h, LGDSP_ERR_INVALID_ARG, "trapezoidal filter does not fit the Savitzky-Golay trace");
    if (p->sg_min_n < 1 || p->sg_max_n < 1 || p->trap_min_n < 1 || p->trap_max_n < 1) return fail(h, LGDSP_ERR_INVALID_ARG, "min_n / max_n must be >= 1");
    if (p->max_triggers < 1 || p->max_triggers > LGDSP_SIPM_MAX_TRIGGERS) return fail(h, LGDSP_ERR_INVALID_ARG, "max_triggers outside 1..%d", LGDSP_SIPM_MAX_TRIGGERS);
    D = SipmDev{};
    D.n = n; D.kind = p->sample_kind; D.t_first = p->t_first_ns; D.dt = p->dt_ns;
    D.trunc_from = p->trunc_from; D.trunc_until = p->trunc_until;
    D.sg_taps = sg.n_taps; D.sg_off = sg.offset;
    for (int k = 0; k < sg.n_taps; ++k) D.sgh[k] = sg.h[k];
    D.sg_min_n = p->sg_min_n; D.sg_max_n = p->sg_max_n;
    D.sg_min_thr = p->sg_min_thr; D.sg_max_thr = p->sg_max_thr; D.sg_nsigma = p->sg_nsigma;
    D.sg_min_dc = p->sg_min_dc; D.sg_max_dc = p->sg_max_dc; D.sg_nsigma_dc = p->sg_nsigma_dc;
    D.ta = t.navg; D.tg = t.ngap; D.ta2 = t.navg2; D.tL = t.navg + t.ngap + t.navg2;
    D.inv1 = 1.0 / t.navg; D.inv2 = 1.0 / t.navg2; D.km1 = p->pz_km1;
    D.trap_min_n = p->trap_min_n; D.trap_max_n = p->trap_max_n;
    D.trap_min_thr = p->trap_min_thr; D.trap_max_thr = p->trap_max_thr; D.trap_nsigma = p->trap_nsigma;
    D.trap_min_dc = p->trap_min_dc; D.trap_max_dc = p->trap_max_dc; D.trap_nsigma_dc = p->trap_nsigma_dc;
    D.cap = p->max_triggers;
    CK(sipm_configure(n, p->sample_kind, &bps));
    if (bps < 1) return fail(h, LGDSP_ERR_CUDA, "sipm kernel does not fit on an SM");
    return LGDSP_OK;
}

int lgdsp_sipm_run_device(lgdsp_handle* h, const lgdsp_sipm_params* p, const void* d_wf, int64_t n_events, int64_t ld_samples,
                          double* d_rows, double* d_trig)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    SipmDev D;
    int bps = 0;
    int rc = sipm_prepare(h, p, D, bps);
    if (rc) return rc;
    if (n_events < 0) return fail(h, LGDSP_ERR_INVALID_ARG, "n_events < 0");
    if (n_events == 0) return LGDSP_OK;
    if (!d_wf || !d_rows || !d_trig) return fail(h, LGDSP_ERR_INVALID_ARG, "NULL pointer");
    if (ld_samples < D.n) return fail(h, LGDSP_ERR_INVALID_ARG, "ld_samples (%lld) < n_samples (%d)", (long long)ld_samples, D.n);
    const long long cap = (long long)h->sm_count * bps;
    const int grid = (int)(n_events < cap ? n_events : cap);
    CK(cudaEventRecord(h->ev0, h->stream));
    sipm_launch(D, d_wf, n_events, ld_samples, d_rows, d_trig, grid, h->stream);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, h->stream));
    h->timed = true;
    h->launches += 1;
    return LGDSP_OK;
}

int lgdsp_sipm_run(lgdsp_handle* h, const lgdsp_sipm_params* p, const void* wf, int64_t n_events, int64_t ld_samples, double* rows,
                   double* trig)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    SipmDev D;
    int bps = 0;
    int rc = sipm_prepare(h, p, D, bps);
    if (rc) return rc;
    if (n_events < 0) return fail(h, LGDSP_ERR_INVALID_ARG, "n_events < 0");
    if (n_events == 0) return LGDSP_OK;
    if (!wf || !rows || !trig) return fail(h, LGDSP_ERR_INVALID_ARG, "NULL pointer");
    if (ld_samples < D.n) return fail(h, LGDSP_ERR_INVALID_ARG, "ld_samples (%lld) < n_samples (%d)", (long long)ld_samples, D.n);
    const size_t sb = (size_t)D.kind;   // bytes per sample
    const size_t per_trig = (size_t)LGDSP_SIPM_NLIST * LGDSP_SIPM_NFIELD * D.cap;
    // double-buffered: the H2D copy of chunk k+1 (copy stream) overlaps the kernel and the D2H copies of chunk k
    const int64_t chunk = n_events < 8192 ? n_events : 8192;
    const size_t out_per = LGDSP_SIPM_NCOL + per_trig;
    rc = ensure_staging(h, (size_t)chunk * D.n * sb, (size_t)2 * chunk * out_per * sizeof(double));
    if (rc) return rc;
    const long long gcap = (long long)h->sm_count * bps;
    const unsigned char* src = static_cast<const unsigned char*>(wf);
    int c = 0;
    for (int64_t e0 = 0; e0 < n_events; e0 += chunk, ++c) {
        const int b = c & 1;
        const int64_t ne = (n_events - e0) < chunk ? (n_events - e0) : chunk;
        if (c >= 2) CK(cudaStreamWaitEvent(h->s_copy, h->ev_free[b], 0));
        CK(cudaMemcpy2DAsync(h->d_in[b], (size_t)D.n * sb, src + (size_t)e0 * ld_samples * sb, (size_t)ld_samples * sb, (size_t)D.n * sb,
                             (size_t)ne, cudaMemcpyHostToDevice, h->s_copy));
        CK(cudaEventRecord(h->ev_ready[b], h->s_copy));
        CK(cudaStreamWaitEvent(h->stream, h->ev_ready[b], 0));
        double* d_rows = h->d_rows + (size_t)b * chunk * out_per;
        double* d_trig = d_rows + (size_t)chunk * LGDSP_SIPM_NCOL;
        const int grid = (int)(ne < gcap ? ne : gcap);
        sipm_launch(D, h->d_in[b], ne, D.n, d_rows, d_trig, grid, h->stream);
        CK(cudaGetLastError());
        h->launches += 1;
        CK(cudaEventRecord(h->ev_free[b], h->stream));
        CK(cudaMemcpyAsync(rows + e0 * LGDSP_SIPM_NCOL, d_rows, (size_t)ne * LGDSP_SIPM_NCOL * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(trig + (size_t)e0 * per_trig, d_trig, (size_t)ne * per_trig * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->s_copy));
    CK(cudaStreamSynchronize(h->stream));
    return LGDSP_OK;
}

int lgdsp_sipm_list_pointers_device(lgdsp_handle* h, const double* d_rows, int64_t n_events, int32_t list, int32_t max_triggers,
                                    int64_t* d_elem_ptr, int64_t* total)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    if (n_events < 0 || list < 0 || list >= LGDSP_SIPM_NLIST || max_triggers < 1) return fail(h, LGDSP_ERR_INVALID_ARG, "bad n_events / list / max_triggers");
    if (!d_elem_ptr || (n_events > 0 && !d_rows)) return fail(h, LGDSP_ERR_INVALID_ARG, "NULL pointer");
    static_assert(sizeof(long long) == sizeof(int64_t), "int64");
    sipm_count_scan_launch(d_rows, n_events, list, max_triggers, reinterpret_cast<long long*>(d_elem_ptr), h->stream);
    CK(cudaGetLastError());
    h->launches += 1;
    long long tot = 0;
    CK(cudaMemcpyAsync(&tot, d_elem_ptr + n_events, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (total) *total = tot;
    return LGDSP_OK;
}

int lgdsp_sipm_list_gather_device(lgdsp_handle* h, const double* d_trig, int64_t n_events, int32_t list, int32_t max_triggers,
                                  const int64_t* d_elem_ptr, double* d_flat, int64_t flat_stride)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    if (n_events < 0 || list < 0 || list >= LGDSP_SIPM_NLIST || max_triggers < 1 || flat_stride < 0) return fail(h, LGDSP_ERR_INVALID_ARG, "bad argument");
    if (n_events == 0) return LGDSP_OK;
    if (!d_trig || !d_elem_ptr || !d_flat) return fail(h, LGDSP_ERR_INVALID_ARG, "NULL pointer");
    sipm_compact_launch(d_trig, n_events, list, max_triggers, reinterpret_cast<const long long*>(d_elem_ptr), d_flat, flat_stride, h->stream);
    CK(cudaGetLastError());
    h->launches += 1;
    return LGDSP_OK;
}

// single-trace primitives (host buffers)
static int prim_run(lgdsp_handle* h, int mode, const double* y, int n, double a, double b, double t0, double dt, int min_n, int max_n,
                    int cap, double* out_host, int n_out, int32_t* n_found)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    if (n < 0 || n > 65536) return fail(h, LGDSP_ERR_UNSUPPORTED, "n = %d: single-trace primitives take 0 .. 65536 samples", n);
    if (n > 0 && !y) return fail(h, LGDSP_ERR_INVALID_ARG, "trace pointer is NULL");
    int rc = ensure_staging(h, (size_t)(n > 0 ? n : 1) * sizeof(double), (size_t)(n_out + 2) * sizeof(double));
    if (rc) return rc;
    if (n > 0) CK(cudaMemcpyAsync(h->d_in[0], y, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    int* d_cnt = reinterpret_cast<int*>(h->d_rows + n_out);
    CK(cudaMemsetAsync(d_cnt, 0, sizeof(int), h->stream));
    sipm_prim_launch(mode, reinterpret_cast<const double*>(h->d_in[0]), n, a, b, t0, dt, min_n, max_n, cap, h->d_rows, d_cnt, h->stream);
    CK(cudaGetLastError());
    h->launches += 1;
    CK(cudaMemcpyAsync(out_host, h->d_rows, (size_t)n_out * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    int cnt = 0;
    CK(cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (n_found) *n_found = cnt;
    return LGDSP_OK;
}

int lgdsp_thresholdstats(lgdsp_handle* h, const double* y, int32_t n, double min, double max, int32_t mad, double* out)
{
    if (!out) return h ? fail(h, LGDSP_ERR_INVALID_ARG, "output pointer is NULL") : LGDSP_ERR_INVALID_ARG;
    return prim_run(h, mad ? 1 : 0, y, n, min, max, 0.0, 1.0, 1, 1, 1, out, 1, nullptr);
}

int lgdsp_intersect_maximum(lgdsp_handle* h, const double* y, int32_t n, double t_first_ns, double dt_ns, double threshold,
                            int32_t min_n, int32_t max_n, int32_t max_triggers, double* x, double* x_high, double* x_tot,
                            double* max, int32_t* n_found)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    if (!x || !x_high || !x_tot || !max || !n_found) return fail(h, LGDSP_ERR_INVALID_ARG, "output pointer is NULL");
    if (min_n < 1 || max_n < 1 || max_triggers < 1 || max_triggers > 65536) return fail(h, LGDSP_ERR_INVALID_ARG, "min_n / max_n / max_triggers out of range");
    std::vector<double> buf((size_t)4 * max_triggers);
    int rc = prim_run(h, 2, y, n, threshold, 0.0, t_first_ns, dt_ns, min_n, max_n, max_triggers, buf.data(), 4 * max_triggers, n_found);
    if (rc) return rc;
    const size_t c = (size_t)max_triggers;
    memcpy(x, buf.data(), c * sizeof(double));
    memcpy(x_high, buf.data() + c, c * sizeof(double));
    memcpy(x_tot, buf.data() + 2 * c, c * sizeof(double));
    memcpy(max, buf.data() + 3 * c, c * sizeof(double));
    return LGDSP_OK;
}

// ---------------------------------------------------------------------------------------------------
// MultiIntersect  (/root/reference/src/multi_intersect.jl:10-121)
// ---------------------------------------------------------------------------------------------------
static int mi_prepare(lgdsp_handle* h, const lgdsp_multi_intersect_params* p, MiDev& D)
{
    if (!p) return fail(h, LGDSP_ERR_INVALID_ARG, "params is NULL");
    if (p->struct_size != sizeof(lgdsp_multi_intersect_params) || p->version != LGDSP_PARAMS_VERSION)
        return fail(h, LGDSP_ERR_INVALID_ARG, "lgdsp_multi_intersect_params: size/version mismatch");
    if (p->n_samples < 2) return fail(h, LGDSP_ERR_INVALID_ARG, "n_samples < 2");
    if (p->n_thresholds < 1 || p->n_thresholds > LGDSP_MI_MAX_THR) return fail(h, LGDSP_ERR_UNSUPPORTED, "n_thresholds outside 1..%d", LGDSP_MI_MAX_THR);
    if (!(p->dt_ns > 0) || !std::isfinite(p->t_first_ns)) return fail(h, LGDSP_ERR_INVALID_ARG, "bad time axis");
    if (p->min_n < 1) return fail(h, LGDSP_ERR_INVALID_ARG, "min_n must be >= 1");
    if (p->half_window < 1 || p->half_window > LGDSP_MI_MAX_HALF || p->degree < 0 || p->degree > LGDSP_MAX_DNI_DEG ||
        p->degree >= 2 * p->half_window || p->rate < 1 || 2 * p->half_window * p->rate > 256)
        return fail(h, LGDSP_ERR_UNSUPPORTED, "polynomial window / degree / sampling rate outside the supported range");
    D = MiDev{};
    D.len = p->n_samples; D.n_thr = p->n_thresholds; D.t0 = p->t_first_ns; D.dt = p->dt_ns;
    D.min_n = p->min_n; D.n = p->half_window; D.degree = p->degree; D.rate = p->rate;
    for (int j = 0; j < p->n_thresholds; ++j) D.ratios[j] = p->ratios[j];
    for (int i = 0; i < 2 * p->half_window * (p->degree + 1); ++i) D.A[i] = p->A[i];
    return LGDSP_OK;
}

int lgdsp_multi_intersect_run_device(lgdsp_handle* h, const lgdsp_multi_intersect_params* p, const double* d_y, int64_t n_events,
                                     int64_t ld_samples, double* d_x, int32_t* d_flags)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    MiDev D;
    int rc = mi_prepare(h, p, D);
    if (rc) return rc;
    if (n_events < 0) return fail(h, LGDSP_ERR_INVALID_ARG, "n_events < 0");
    if (n_events == 0) return LGDSP_OK;
    if (!d_y || !d_x || !d_flags) return fail(h, LGDSP_ERR_INVALID_ARG, "NULL pointer");
    if (ld_samples < D.len) return fail(h, LGDSP_ERR_INVALID_ARG, "ld_samples < n_samples");
    CK(cudaEventRecord(h->ev0, h->stream));
    multi_intersect_launch(D, d_y, n_events, ld_samples, d_x, d_flags, h->stream);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, h->stream));
    h->timed = true;
    h->launches += 1;
    return LGDSP_OK;
}

int lgdsp_multi_intersect_run(lgdsp_handle* h, const lgdsp_multi_intersect_params* p, const double* y, int64_t n_events,
                              int64_t ld_samples, double* x, int32_t* flags)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    MiDev D;
    int rc = mi_prepare(h, p, D);
    if (rc) return rc;
    if (n_events < 0) return fail(h, LGDSP_ERR_INVALID_ARG, "n_events < 0");
    if (n_events == 0) return LGDSP_OK;
    if (!y || !x || !flags) return fail(h, LGDSP_ERR_INVALID_ARG, "NULL pointer");
    if (ld_samples < D.len) return fail(h, LGDSP_ERR_INVALID_ARG, "ld_samples < n_samples");
    const int64_t chunk = n_events < 4096 ? n_events : 4096;
    const size_t out_per = (size_t)D.n_thr * sizeof(double) + sizeof(int32_t);
    rc = ensure_staging(h, (size_t)chunk * D.len * sizeof(double), (size_t)chunk * out_per + 16);
    if (rc) return rc;
    double* d_x = h->d_rows;
    int* d_f = reinterpret_cast<int*>(h->d_rows + (size_t)chunk * D.n_thr);
    for (int64_t e0 = 0; e0 < n_events; e0 += chunk) {
        const int64_t ne = (n_events - e0) < chunk ? (n_events - e0) : chunk;
        CK(cudaMemcpy2DAsync(h->d_in[0], (size_t)D.len * sizeof(double), y + e0 * ld_samples, (size_t)ld_samples * sizeof(double),
                             (size_t)D.len * sizeof(double), (size_t)ne, cudaMemcpyHostToDevice, h->stream));
        multi_intersect_launch(D, reinterpret_cast<const double*>(h->d_in[0]), ne, D.len, d_x, d_f, h->stream);
        CK(cudaGetLastError());
        h->launches += 1;
        CK(cudaMemcpyAsync(x + e0 * D.n_thr, d_x, (size_t)ne * D.n_thr * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(flags + e0, d_f, (size_t)ne * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    return LGDSP_OK;
}

// ---------------------------------------------------------------------------------------------------
// synthetic input
// ---------------------------------------------------------------------------------------------------
int lgdsp_synth_generate_device(lgdsp_handle* h, const lgdsp_synth_params* sp, int64_t first_event, int64_t n_events,
                                int64_t ld_samples, uint16_t* d_wf)
{
    if (!h) return LGDSP_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    if (!sp || sp->n_samples < 8 || sp->n_samples % 8 != 0 || sp->n_samples > LGDSP_MAX_SAMPLES || !(sp->tau_samples > 0))
        return fail(h, LGDSP_ERR_INVALID_ARG, "bad synth params");
    int rc = check_wf(h, d_wf, n_events, ld_samples, sp->n_samples, true);
    if (rc) return rc;
    if (n_events == 0) return LGDSP_OK;
    synth_launch(*sp, first_event, n_events, ld_samples, d_wf, h->stream);
    CK(cudaGetLastError());
    h->launches += 1;
    return LGDSP_OK;
}

}  // extern "C"
