// fine_doppler.cuh -- launch interface of the fine-Doppler kernels (fine_doppler.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gb {

struct FineArgs {
    const float2* x;               // sample ring / uploaded recording
    unsigned long long start;      // absolute index of long_samples[0]
    unsigned long long mask;       // ring mask (all ones for a linear buffer)
    unsigned use;                  // size_signal_use = (long_ms - 1) * samples per code
    float fs;
    int log_a, log_b;              // P2 = next_power_of_two(use) = 2^(log_a + log_b)
    const int8_t* codes;           // n_req x 1023 chips (+-1)
    const unsigned long long* code_phase;  // n_req
    float2* mean;                  // 1 (written by fine_mean_kernel)
    float2* Y;                     // n_req x 8 x P2 scratch (pass 1 -> pass 2, L2-resident)
    unsigned long long* best;      // n_req packed (magnitude bits << 32 | ~index)
    float* mag_out;                // optional n_req x 8*P2 magnitudes (diagnostics), or nullptr
};

size_t fine_smem_bytes(int log_a, int log_b);
// mean -> column pass -> row pass + arg-max, all on stream st
cudaError_t fine_launch(const FineArgs& a, int n_req, const float2* x, unsigned long long start, unsigned long long mask,
                        unsigned long long n_long, cudaStream_t st);

}  // namespace gb
