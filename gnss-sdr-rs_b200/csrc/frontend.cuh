// frontend.cuh -- table-driven digital front-end (frontend.cu): DigitalFrontend::process_block (rf/frontend.rs:32-62)
// with the sequential f32 NCO phase accumulator replaced by its own, precomputed, eventually periodic orbit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace gb {

// The reference's NCO state is one f32 advanced by `acc = (acc + step) % 2048.0` per sample (frontend.rs:48-52): a
// deterministic map on a finite set, so the sequence from acc = 0 is a tail of `mu` states followed by a cycle of
// `lambda` states (for the steps met in practice mu <= 3 and lambda <= 2^23: every wrap passes through the [2048, 4096)
// binade and leaves a multiple of 2^-12).  fe_build_phase_orbit finds (mu, lambda) with Brent's algorithm and returns
// the phases of the tail followed by `reps` copies of the cycle, reps chosen so that the periodic part is at least
// `min_period` long (the kernel wraps a position with ONE conditional subtraction of the padded period).
// Returns false if no cycle of at most `cap` states exists (the caller falls back to the sequential kernel).
bool fe_build_phase_orbit(float step, uint64_t cap, uint64_t min_period, std::vector<float>& phase, uint64_t* mu,
                          uint64_t* period);
// LUT index of a phase: `phase_accumulator as usize % LUT_SIZE` (saturating cast, frontend.rs:49)
uint16_t fe_lut_index(float phase);
// position in the orbit table of the sample with absolute number `count` (samples since configure)
uint64_t fe_orbit_pos(uint64_t count, uint64_t mu, uint64_t period);

// One launch = one process_block call of n samples (n % 8 == 0) appended to the ring at `head`.
// idx_tab: orbit table of LUT indices (tail + padded cycle), pos0 = fe_orbit_pos(samples before this call).
// bias: 16 floats on the device (bias_re[8], bias_im[8]), updated in place.
cudaError_t fe_launch_table(const float2* src, float2* ring, unsigned long long head, unsigned long long mask,
                            unsigned long long n, const float* lut, float* bias, const uint16_t* idx_tab,
                            unsigned long long pos0, unsigned long long mu, unsigned long long period, float alpha,
                            float con, cudaStream_t st);

// Tolerance mode (frontend.cu, "segmented scan"): the same call as three multi-CTA launches.  scratch: at least
// fe_parallel_scratch(n) float2 elements on the device.
size_t fe_parallel_scratch(unsigned long long n);
cudaError_t fe_launch_parallel(const float2* src, float2* ring, unsigned long long head, unsigned long long mask,
                               unsigned long long n, const float* lut, float* bias, const uint16_t* idx_tab,
                               unsigned long long pos0, unsigned long long mu, unsigned long long period, float alpha,
                               float con, float2* scratch, cudaStream_t st);

}  // namespace gb
