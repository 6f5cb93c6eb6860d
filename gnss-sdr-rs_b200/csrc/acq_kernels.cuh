// acq_kernels.cuh -- launch interface of the fused acquisition kernels (acq_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gnss_b200.h"

namespace gb {

// explicit tuning switches (gb_tuning_set, include/gnss_b200.h); defined in gnss_b200.cu
int tuning(const char* key, int dflt);

struct AcqArgs {
    const float2* iq;        // sample ring / chunk base
    unsigned long long iq_start;  // absolute index of sample 0 of block 0
    unsigned long long iq_mask;   // ring mask (all ones for a linear chunk)
    const float2* tables;    // D x N wipe-off tables (DopplerShiftTable::table)
    const float2* code_fft;  // n_prn x N code spectra, scrambled + transposed for the middle stage
    const float2* tw;        // per-stage twiddles W_L^(i q), layout [stage][q-1][i]
    const float2* rot;       // D x n_coh coherent rotators or nullptr
    const int* rows;         // active PRN rows (index into code_fft / cells rows)
    int D, K, n_coh, n_active;
    int spc;                 // samples per chip for the two-peak exclusion (0 = off)
    gb_acq_cell* cells;      // n_prn x D
    float* row_out;          // diagnostics: accumulated power row of (rows[0], d0)
    int d0;                  // Doppler bin for row_out
    float2* spec;            // shared-forward chain: n_d x n_groups x N scrambled spectra (scratch)
    int d_lo;                // shared-forward chain: first Doppler bin of this slab
    int g_lo, g_cnt;         // shared-forward chain: groups covered by this forward launch (g_cnt == 0: none)
    const int* npos;         // prime-factor plans: code-phase index n(l) of line position l (else unused)
    const float2* otw;       // cluster plans: outer twiddles W_N^(i q), layout [q-1][i]
    float* acc_rows;         // cluster plans: (n_active*D) x N accumulated power rows (scratch)
    // Doppler aliasing (shared-forward chain): bins whose carriers differ by m * fs/N share ONE forward spectrum
    // (X_d[k] = X_base[k + m]); the inverse kernel pairs it with the code spectrum shifted the other way.
    const int* fwd_bins;     // forward launch: bin computed by slot dl (nullptr: d_lo + dl)
    const int2* inv_map;     // inverse launch: per bin {spectrum slot, index of the shifted code-spectrum set} (nullptr: {dl, 0})
    int n_prn;               // rows per code-spectrum set
    int plain_inverse;       // shared-forward chain: 1 = acq_inverse_kernel even where a leftover-warp form exists (A/B)
    // tensor-pipe A/B of the N = 4092 inverse kernel (gb_tuning_set("acq_tc", 1), acq_lw.cu): scratch for the spectra / code
    // spectra in fragment order (nullptr = off), forward slots in a.spec when aliased (0: the slab's n_d), code-spectrum
    // sets behind a.code_fft, 1 = code_tc already holds them
    float2* spec_tc;
    float2* code_tc;
    int tc_n_fwd, tc_n_code_sets, tc_code_fresh;
};

struct FftArgs {
    const void* in;          // float2 (or float if real_in) batch x n, natural order
    void* out;               // float2 (or float if power_out) batch x n_out
    const float2* tw;
    const int* freq_of_pos;  // scrambled position -> natural frequency index
    const int* npos;         // prime-factor plans: input index of line position l (else unused)
    int real_in, power_out, n_out;
};

int acq_plan_index(int n);                    // -1 if there is no plan for n
int acq_plan_sizes(int* sizes, int cap);      // list of planned sizes
int acq_plan_radices(int plan, int* radices); // returns number of stages
int acq_plan_threads(int plan);
int acq_plan_spec_len(int plan);             // complex elements of one stored spectrum (rows padded to 128 B; >= n)
int acq_plan_spec_stride(int plan);          // row stride of the stored [q][b] layout
int acq_plan_twiddles(int plan);             // length of the per-stage twiddle buffer ([stage][q-1][i])
int acq_plan_supports_alias(int plan);       // 1: the shared chain has a Doppler-aliasing inverse kernel for this plan
int acq_plan_is_pfa(int plan);               // 1: Good-Thomas prime-factor plan (inputs in line order, see PfaPlan)
size_t acq_plan_smem(int plan);

cudaError_t acq_launch_search(int plan, const AcqArgs& a, cudaStream_t st);
// forward kernel over n_d bins [a.d_lo, a.d_lo+n_d) then inverse kernel over n_active x n_d cells
cudaError_t acq_launch_shared(int plan, const AcqArgs& a, int n_d, cudaStream_t st);
// N = 4092 shared chain: inverse kernel with a leftover warp (acq_lw.cu); the forward spectra must be in a.spec
cudaError_t acq_launch_inverse_lw4092(const AcqArgs& a, int n_d, cudaStream_t st);
int acq_tc_spec_len();   // complex elements of one spectrum in the fragment order of the tensor-pipe A/B
cudaError_t acq_launch_row(int plan, const AcqArgs& a, cudaStream_t st);
cudaError_t acq_launch_code_fft(int plan, const int8_t* codes, int n_prn, float2* code_fft, const float2* tw,
                                const int* npos, cudaStream_t st);
// prime-factor plans: dst[b*n + l] = src[(start + b*n + npos[l]) & mask] for b0 <= b < b0 + n_blocks
cudaError_t acq_launch_permute(const float2* src, unsigned long long start, unsigned long long mask, const int* npos, int n,
                               int b0, int n_blocks, float2* dst, cudaStream_t st);
// forward kernel alone over n_d bins x groups [a.g_lo, a.g_lo + a.g_cnt)
cudaError_t acq_launch_forward(int plan, const AcqArgs& a, int n_d, cudaStream_t st);
cudaError_t acq_launch_fft(int plan, int inverse, const FftArgs& a, int batch, cudaStream_t st);
// steps_dev[d] = 2*pi*(f_if+f_d)/fs (f32, host-evaluated in the reference's order)
// dst[(s * n_prn + p) * n + t] = src[p * n + gidx[s * n + t]]: the code spectra re-indexed for n_shift frequency shifts
cudaError_t acq_launch_shift_codes(const float2* src, const int* gidx, int n_shift, int n_prn, int n, float2* dst, cudaStream_t st);
cudaError_t acq_launch_doppler_tables(const float* steps_dev, int D, int n, float2* tables, cudaStream_t st);

// cluster plans (code period larger than one CTA's shared memory): radix-RO outer stage over a thread-block cluster
int acq_cluster_supported(int n);
int acq_cluster_inner(int n);   // inner plan length (has a regular plan)
int acq_cluster_outer(int n);   // outer radix = cluster size
cudaError_t acq_cluster_launch_search(const AcqArgs& a, cudaStream_t st);
cudaError_t acq_cluster_launch_code_fft(const int8_t* codes, int n_prn, float2* code_fft, const float2* tw, const float2* otw,
                                        cudaStream_t st);

}  // namespace gb
