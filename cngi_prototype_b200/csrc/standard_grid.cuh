// standard_grid.cuh -- parameter block and packed-pair arithmetic shared by the standard-gridder kernels
// (standard_grid.cu: naive + register-window "track" kernel; standard_grid_shift.cu: shifting-window kernel).
#pragma once
#include "common.cuh"

namespace cngi {

struct StdParams {
    int n_time, n_baseline, n_chan, n_pol;
    int n_ic, n_ip, n_u, n_v;
    const void *vis;
    const void *weight;
    const uint8_t *flag;
    const double *uvw;
    const double *freq;
    const int64_t *chan_map;
    const int64_t *pol_map;
    const double *cgk;
    void *grid;
    double *sum_weight;
    void *psf_grid;          // fused image + psf pass only (real, same plane layout as grid)
    double *psf_sum_weight;
    double dl, dm;
    int support, oversampling, do_psf, chan_mode;
    int table_len;
    // track kernel decomposition
    int G, log2G, seg_len, n_seg, n_cspan, n_pgrp;
    int c_lo, c_n;         // channel window handled by this launch (bounds the shared-memory channel table)
    long long n_tasks;
    const double *scale;   // naive kernel: [2, n_chan] uv_scale table
    // fused imaging weights (window kernel, IWF): the weight handed in is the NATURAL weight; the imaging weight of
    // _standard_imaging_weight_degrid_jit (_standard_grid.py:466-518) is formed per sample in phase 1 from the density grid
    const double *iw_density;                        // element strides below: (u, v, imaging chan, imaging pol)
    long long iw_ds_u, iw_ds_v, iw_ds_c, iw_ds_p;
    const double *iw_bf;                             // Briggs factors [2, n_ic, n_ip], contiguous
    void *iw_out;                                    // optional: imaging weights written like A4 would (sample layout)
    int iw_n_u, iw_n_v, iw_own_scale, iw_pol_shared;                // geometry of the density grid (make_imaging_weight does not pad)
    double iw_dl, iw_dm;
};

__device__ __forceinline__ int chan_of(const StdParams &p, int c)
{
    if (p.chan_mode == CNGI_CHAN_CUBE) return c;
    if (p.chan_mode == CNGI_CHAN_CONTINUUM) return 0;
    return (int)p.chan_map[c];
}
__device__ __forceinline__ int pol_of(const StdParams &p, int ip) { return p.pol_map ? (int)p.pol_map[ip] : ip; }

// Packed pair arithmetic.  Accumulators are (re, im) pairs (complex grid) or (pol0, pol1) pairs (real grid), so
// that on sm_100a every fp32 update is one FFMA2 (fma.rn.f32x2: two FMAs per issue slot, the only way to reach
// the full FP32 rate with three register operands).  fp64 uses two DFMAs.
template <typename T> struct Pair;
template <> struct Pair<float> { using type = float2; };
template <> struct Pair<double> { using type = double2; };

__device__ __forceinline__ void pk_fma_acc(float2 &c, float2 a, float s)   // c += a * (s, s); accumulator tied in place
{
    float2 b = make_float2(s, s);
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(*reinterpret_cast<unsigned long long *>(&c))
        : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
}
__device__ __forceinline__ float2 pk_mul(float2 a, float s)
{
    float2 b = make_float2(s, s), d;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(*reinterpret_cast<unsigned long long *>(&d))
        : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
    return d;
}
__device__ __forceinline__ void pk_fma_acc(double2 &c, double2 a, double s)
{
    c.x = fma(a.x, s, c.x);
    c.y = fma(a.y, s, c.y);
}
__device__ __forceinline__ double2 pk_mul(double2 a, double s) { return make_double2(a.x * s, a.y * s); }

// standard_grid_shift.cu
bool shift_kernel_supported(const cngi_std_grid_args *a, int table_len);
int launch_shift(StdParams p, const cngi_std_grid_args *a, cudaStream_t st);
// standard_grid_window.cu
bool window_kernel_supported(const cngi_std_grid_args *a, int table_len);
int launch_window(StdParams p, const cngi_std_grid_args *a, cudaStream_t st);
int launch_window_dual(StdParams p, const cngi_std_grid_args *a, cudaStream_t st);
int launch_window_iw(StdParams p, const cngi_std_grid_args *a, cudaStream_t st);
// standard_grid_window16_f32.cu / _f64.cu: supports 9 / 11 / 13 / 15
int launch_window16_f32(StdParams p, const cngi_std_grid_args *a, cudaStream_t st);
int launch_window16_f64(StdParams p, const cngi_std_grid_args *a, cudaStream_t st);

}  // namespace cngi
