// trk_kernels.cuh -- launch interface of the batched E/P/L correlator + loop kernels (trk_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gnss_b200.h"

namespace gb {

struct TrkArgs {
    const float2* samples;          // device ring (or linear buffer when offsets != nullptr)
    unsigned long long mask;        // ring mask (all ones for a linear buffer)
    unsigned long long head;        // absolute index one past the newest sample
    unsigned long long capacity;    // ring capacity (0 = unbounded linear buffer)
    const unsigned long long* offsets;  // correlate-only: per-channel start of its samples
    gb_trk_channel* ch;             // n_channels channel states (device)
    const int8_t* ca_table;         // 32 x 1023 chips (device)
    int n_channels, n_epochs, filters, n_max;
    gb_trk_corr* corr;              // last epoch's six sums per channel
    float* prompt_hist;             // n_epochs x n_channels x 2 or nullptr
    uint8_t* ran;                   // per channel: epochs consumed in this launch (saturating at 255)
    uint8_t* lost;                  // per channel: SatelliteLost emitted
    int dbg = 0;                    // tuning bits (gb_tuning_set("trk_dbg", v)): 1 = no L2 prefetch of the window after next
};

cudaError_t trk_launch(const TrkArgs& a, int mode, cudaStream_t st);
// dynamic shared memory of the ORDERED kernel for epochs of up to n_max samples
size_t trk_ordered_smem_bytes(int n_max);
// warp-specialised FAST kernel for ring-fed epochs with loop filters (trk_ws.cu)
bool trk_ws_supported(const TrkArgs& a);
cudaError_t trk_ws_launch(const TrkArgs& a, cudaStream_t st, int variant);

// explicit tuning switches (gb_tuning_set in include/gnss_b200.h; there are no environment-variable switches)
int tuning(const char* key, int dflt);

}  // namespace gb
