// Host-callable launchers of the CUDA kernels (internal to the library).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "lgdsp_device.cuh"

namespace lgdsp {

// fused dsp_icpc kernel
int icpc_smem_bytes();
int icpc_threads();
cudaError_t icpc_configure(int* max_blocks_per_sm);
void icpc_launch(const IcpcDev& P, const uint16_t* d_wf, long long n_events, long long ld, double* d_rows, int grid,
                 cudaStream_t stream);

// trapezoid sweep kernel
struct SweepVar {
    TrapDev t;
    double pick_ns;
    int mode, pad_;
};
struct SweepDev {
    int n, tx_min_n;
    double t_first, dt;
    int bl_from, bl_until;
    double km1;
    DniDev sig_dni;
    const double* dni_A;     // [LGDSP_MAX_DNI*4]
    const SweepVar* vars;    // device array
    int nvar, pad_;
};
cudaError_t sweep_configure(int* max_blocks_per_sm);
void sweep_launch(const SweepDev& P, const uint16_t* d_wf, long long n_events, long long ld, float* d_out, int grid,
                  cudaStream_t stream);

// synthetic generator
void synth_launch(const lgdsp_synth_params& sp, long long first_event, long long n_events, long long ld, uint16_t* d_wf,
                  cudaStream_t stream);

}  // namespace lgdsp
