// acq_generic.cuh -- any-length fallback plan (acq_generic.cu): Bluestein over a power-of-two Stockham FFT.
#pragma once
#include "acq_kernels.cuh"

namespace gb {

struct GenericPlan;   // chirp + kernel spectra of one length N (f32; the f64 twin is built on first use)

cudaError_t generic_plan_create(int n, GenericPlan** out, cudaStream_t st);
void generic_plan_destroy(GenericPlan* p);
int generic_plan_m(const GenericPlan* p);   // the power-of-two convolution length (N itself for the f32/f64 facade of a power of two)

// batch (<= 32768) of natural-order, unnormalised length-N DFTs; s0, s1: scratch of batch x generic_plan_m() elements each
cudaError_t generic_dft_f32(GenericPlan* p, int inverse, const float2* in, float2* out, int batch, float2* s0, float2* s1, cudaStream_t st);
cudaError_t generic_dft_f64(GenericPlan* p, int inverse, const double2* in, double2* out, int batch, double2* s0, double2* s1,
                            cudaStream_t st);
cudaError_t generic_real_to_complex(const float* in, float2* out, size_t total, cudaStream_t st);
cudaError_t generic_r2c_f64(const double* in, double2* out, size_t total, cudaStream_t st);
// out[b][k] = x[b][k] (power = 0) or |x[b][k]|^2 (power = 1) for k < n_out <= n
cudaError_t generic_take_f32(const float2* x, void* out, int n, int n_out, int batch, int power, cudaStream_t st);
cudaError_t generic_take_f64(const double2* x, void* out, int n, int n_out, int batch, int power, cudaStream_t st);

// AcquisitionWorker::new: n_prn natural-order code spectra
cudaError_t generic_code_fft(GenericPlan* p, const int8_t* codes_dev, int n_prn, float2* code_fft, float2* s0, float2* s1, cudaStream_t st);
// search_satellite over Doppler bins [d_lo, d_lo + n_d): acc[(row_index * D + d) * N + n] = accumulated power;
// s0, s1: scratch of a.n_active * n_d * generic_plan_m() complex each (a.n_active * n_d <= 32768)
cudaError_t generic_search_slab(GenericPlan* p, const AcqArgs& a, int d_lo, int n_d, float* acc, float2* s0, float2* s1, cudaStream_t st);

// accumulated power rows -> cells (acq_cluster.cu): rows_total = n_active * D rows of n floats
cudaError_t acq_launch_reduce_rows(const float* acc_rows, int n, int D, int n_active, const int* rows, int spc, gb_acq_cell* cells,
                                   cudaStream_t st);

}  // namespace gb
