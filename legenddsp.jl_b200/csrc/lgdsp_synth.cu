// Synthetic waveform generator kernels (device) and the identical host loop.
#include <cuda_runtime.h>
#include "lgdsp_kernels.h"
#include "lgdsp_synth.cuh"

namespace lgdsp {

// one thread per group of 8 samples (one 16-byte store)
__global__ void synth_kernel(lgdsp_synth_params sp, long long first_event, long long n_events, long long ld,
                             uint16_t* __restrict__ wf)
{
    const int groups = sp.n_samples / 8;
    const long long total = n_events * (long long)groups;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long le = t / groups;
        const int g = (int)(t - le * groups);
        const long long e = first_event + le;
        const lgdsp_synth::EventPars ep = lgdsp_synth::event_pars(sp, e);
        uint16_t v[8];
        lgdsp_synth::sample_group(sp, ep, e, g * 8, v);
        lgdsp_synth::sample_group(sp, ep, e, g * 8 + 4, v + 4);
        uint4 o;
        o.x = v[0] | ((uint32_t)v[1] << 16);
        o.y = v[2] | ((uint32_t)v[3] << 16);
        o.z = v[4] | ((uint32_t)v[5] << 16);
        o.w = v[6] | ((uint32_t)v[7] << 16);
        *reinterpret_cast<uint4*>(wf + le * ld + g * 8) = o;
    }
}

void synth_launch(const lgdsp_synth_params& sp, long long first_event, long long n_events, long long ld, uint16_t* d_wf,
                  cudaStream_t stream)
{
    const long long total = n_events * (long long)(sp.n_samples / 8);
    if (total <= 0) return;
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    synth_kernel<<<(unsigned)blocks, 256, 0, stream>>>(sp, first_event, n_events, ld, d_wf);
}

}  // namespace lgdsp

extern "C" int lgdsp_synth_generate_host(const lgdsp_synth_params* sp, int64_t first_event, int64_t n_events,
                                         int64_t ld_samples, uint16_t* wf)
{
    if (!sp || !wf || sp->n_samples <= 0 || sp->n_samples % 8 != 0 || ld_samples < sp->n_samples || n_events < 0)
        return LGDSP_ERR_INVALID_ARG;
    for (int64_t le = 0; le < n_events; ++le) {
        const int64_t e = first_event + le;
        const lgdsp_synth::EventPars ep = lgdsp_synth::event_pars(*sp, e);
        for (int i0 = 0; i0 < sp->n_samples; i0 += 4) lgdsp_synth::sample_group(*sp, ep, e, i0, wf + le * ld_samples + i0);
    }
    return LGDSP_OK;
}
