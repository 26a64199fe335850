// common.cuh -- shared device/host helpers for libcngi_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cmath>
#include "../../include/cngi_b200.h"

namespace cngi {

// ---- error plumbing (thread-local message, int status; the C ABI never throws) -----------------
void set_error(const char *fmt, ...);
#define CNGI_CUDA_TRY(expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            ::cngi::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                              __LINE__);                                                         \
            return CNGI_ERR_CUDA;                                                                \
        }                                                                                        \
    } while (0)
#define CNGI_REQUIRE(cond, ...)                  \
    do {                                         \
        if (!(cond)) {                           \
            ::cngi::set_error(__VA_ARGS__);      \
            return CNGI_ERR_INVALID;             \
        }                                        \
    } while (0)

constexpr double kSpeedOfLight = 299792458.0;

template <typename T> struct Cplx;
template <> struct Cplx<float> { using type = float2; };
template <> struct Cplx<double> { using type = double2; };

// ---- bit-exact index math ---------------------------------------------------------------------
// uv_scale = -(freq * delta * n) / c, evaluated left to right like numpy does at
// _standard_grid.py:275-276.  The _rn intrinsics are IEEE-754 correctly rounded and are never
// contracted into FMAs, so the result equals the host/numba value bit for bit.
__device__ __forceinline__ double uv_scale_of(double freq, double delta, int n)
{
    return -__ddiv_rn(__dmul_rn(__dmul_rn(freq, delta), (double)n), kSpeedOfLight);
}

struct CellPos {
    int uc, vc;      // centre cell (int(x + 0.5), truncation toward zero)
    double u_pos, v_pos;
};

// Restates _standard_grid.py:299-315: returns false when u or v is NaN.
__device__ __forceinline__ bool locate_centre(double uvw_u, double uvw_v, double us, double vs, int n_u, int n_v,
                                              CellPos &c)
{
    const double u = __dmul_rn(uvw_u, us);
    const double v = __dmul_rn(uvw_v, vs);
    if (isnan(u) || isnan(v)) return false;
    c.u_pos = __dadd_rn(u, (double)(n_u / 2));
    c.v_pos = __dadd_rn(v, (double)(n_v / 2));
    c.uc = __double2int_rz(__dadd_rn(c.u_pos, 0.5));
    c.vc = __double2int_rz(__dadd_rn(c.v_pos, 0.5));
    return true;
}

// floor((centre - pos) * oversampling + 0.5)   (_standard_grid.py:321-324)
__device__ __forceinline__ int oversample_offset(int centre, double pos, int oversampling)
{
    const double off = __dsub_rn((double)centre, pos);
    return __double2int_rd(__dadd_rn(__dmul_rn(off, (double)oversampling), 0.5));
}

// Written without `centre +- half` so that a saturated centre (uvw = +-inf or huge: __double2int_rz clamps to INT_MAX /
// INT_MIN) cannot wrap around and pass the test -- such samples are skipped like every other out-of-grid sample.
__device__ __forceinline__ bool stamp_inside(int uc, int vc, int half, int n_u, int n_v)
{
    return (uc < n_u - half) && (vc < n_v - half) && (uc >= half) && (vc >= half);
}

// weighted_data = vis * weight as numba evaluates complex128 * float64 (weight promoted to w + 0j).
// Kept literal so that NaN/Inf masking (_standard_grid.py:340) is bit-exact.
__device__ __forceinline__ void weighted_vis(double a, double b, double w, double &re, double &im)
{
    re = __dsub_rn(__dmul_rn(a, w), __dmul_rn(b, 0.0));
    im = __dadd_rn(__dmul_rn(a, 0.0), __dmul_rn(b, w));
}

__device__ __forceinline__ bool masked(double re, double im)
{
    return isnan(re) || isnan(im) || (re == 0.0 && im == 0.0);
}

// ---- global reductions (native REDG on sm_100a; shared-memory fp atomics are CAS loops) -----------
__device__ __forceinline__ void red_add(float *p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(double *p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(float2 *p, float2 v) { atomicAdd(p, v); }   // REDG.E.ADD.F32x2
__device__ __forceinline__ void red_add(double2 *p, double2 v)
{
    atomicAdd(&p->x, v.x);
    atomicAdd(&p->y, v.y);
}

// ---- packed pair arithmetic: acc(x, y) += a(x, y) * s.  fp32: one FFMA2 (fma.rn.f32x2), fp64: two DFMAs ----------------
__device__ __forceinline__ void pair_fma(float2 &acc, float2 a, float s)
{
    float2 b = make_float2(s, s);
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(*reinterpret_cast<unsigned long long *>(&acc))
        : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
}
__device__ __forceinline__ void pair_fma(double2 &acc, double2 a, double s)
{
    acc.x = fma(a.x, s, acc.x);
    acc.y = fma(a.y, s, acc.y);
}

// Adds `val` into base[index] with one reduction per distinct index in the warp when the warp hits at
// most two distinct indices (continuum imaging: every lane hits the same sum_weight slot).
__device__ __forceinline__ void warp_grouped_add(double *base, int index, double val, bool active)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    unsigned todo = __ballot_sync(FULL, active);
    for (int round = 0; round < 2 && todo; ++round) {
        const int leader = __ffs(todo) - 1;
        const int idx0 = __shfl_sync(FULL, index, leader);
        const bool mine = active && (index == idx0) && ((todo >> lane) & 1u);
        const unsigned grp = __ballot_sync(FULL, mine);
        double v = mine ? val : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == leader) atomicAdd(base + idx0, v);
        todo &= ~grp;
    }
    if ((todo >> lane) & 1u) atomicAdd(base + index, val);
}

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

int sm_count();
int tune_pool_once();   // keep stream-ordered scratch memory in the pool across synchronisations

// uv_scale[0][c], uv_scale[1][c] for all channels in a stream-ordered scratch buffer (one ddiv per channel
// instead of one per sample for the kernels that map a thread to a single sample).  Free with cudaFreeAsync.
int make_uv_scale_table(const double *freq, int n_chan, double dl, double dm, int n_u, int n_v, cudaStream_t st,
                        double **table);

}  // namespace cngi
