// direction_rotate.cu -- N3: uvw rotation + visibility phasor, the step in front of the
// mosaic gridders.  Replaces ngcasa/imaging/direction_rotate.py:190-248 (apply_rotation_matrix, apply_phasor).
//
//   uvw_rot[t,b,:] = uvw[t,b,:] @ R[field(t)]                                   (:205)
//   d[t,b]         = uvw_rot[t,b,0:end] . phase_rotation[field(t), 0:end]       (:229), end = 2 (common tangent) or 3
//   vis_rot        = vis * exp(i * ((2 pi d) f_c) (1/c0))                       (:239-241)
//
// The phase is formed in the reference's own order -- numpy evaluates `2.0*1j*np.pi*d*f/c` left to right and its
// complex/real division multiplies by the reciprocal -- with never-contracted IEEE operations, so the argument of
// the sin/cos is the reference's bit for bit; a phase of 1e3 rad would otherwise carry 1e-13 of difference per ulp.
// One thread per (t, b, chan) moves all pols with 128-bit accesses: the kernel streams vis in and out once
// (HBM bound; the fp64 sincos runs under the loads).
#include "common.cuh"

namespace cngi {

constexpr long long kIntNan = -2147483648LL;   // cngi/_utils/_constants.py:19

// One warp per integration: the field must be constant over baseline (the reference asserts it, :200,:227).
// idx[0][t] uses ids > -1 (apply_rotation_matrix :198), idx[1][t] ids != INT_NAN (apply_phasor :226).
__global__ void dr_field_index_kernel(const int64_t *__restrict__ field, int n_time, int n_baseline,
                                      const int64_t *__restrict__ rot_field_id, int n_field, int *__restrict__ idx,
                                      int *__restrict__ status)
{
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= n_time) return;
    const int64_t *row = field + (size_t)t * n_baseline;
    for (int which = 0; which < 2; ++which) {
        long long lo = LLONG_MAX, hi = LLONG_MIN;
        for (int b = lane; b < n_baseline; b += 32) {
            const long long f = row[b];
            const bool use = which == 0 ? (f > -1) : (f != kIntNan);
            if (use) {
                lo = min(lo, f);
                hi = max(hi, f);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) {
            int found = -1;
            if (lo == hi)
                for (int i = 0; i < n_field; ++i)
                    if (rot_field_id[i] == lo) {
                        found = i;
                        break;
                    }
            idx[which * n_time + t] = found;
            if (found < 0 && status) atomicExch(status, 1);
        }
    }
}

// uvw[t,b,:] @ R : term order (u R0k + v R1k) + w R2k, no contraction.
__device__ __forceinline__ void rotate_uvw(const double *__restrict__ p, const double *__restrict__ R, double &u,
                                           double &v, double &w)
{
    const double a = p[0], b = p[1], c = p[2];
    u = __dadd_rn(__dadd_rn(__dmul_rn(a, R[0]), __dmul_rn(b, R[3])), __dmul_rn(c, R[6]));
    v = __dadd_rn(__dadd_rn(__dmul_rn(a, R[1]), __dmul_rn(b, R[4])), __dmul_rn(c, R[7]));
    w = __dadd_rn(__dadd_rn(__dmul_rn(a, R[2]), __dmul_rn(b, R[5])), __dmul_rn(c, R[8]));
}

__global__ void dr_uvw_kernel(const double *uvw, double *uvw_rot, const int *__restrict__ idx, int n_baseline,
                              long long n_rows, const double *__restrict__ rotmat)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int f = idx[r / n_baseline];
    double u, v, w;
    if (f < 0) {
        u = v = w = 0.0;   // the reference raises here; status was set by dr_field_index_kernel
    } else {
        rotate_uvw(uvw + 3 * r, rotmat + 9 * f, u, v, w);
    }
    uvw_rot[3 * r + 0] = u;
    uvw_rot[3 * r + 1] = v;
    uvw_rot[3 * r + 2] = w;
}

template <typename T> struct VisIO;
template <> struct VisIO<float> {   // complex64: 2 pols = one 128-bit access
    static __device__ __forceinline__ void load2(const void *p, size_t i, double *re, double *im)
    {
        const float4 x = __ldcs(reinterpret_cast<const float4 *>(p) + i);
        re[0] = x.x; im[0] = x.y; re[1] = x.z; im[1] = x.w;
    }
    static __device__ __forceinline__ void store2(void *p, size_t i, const double *re, const double *im, bool)
    {
        __stcs(reinterpret_cast<float4 *>(p) + i,
               make_float4(__double2float_rn(re[0]), __double2float_rn(im[0]), __double2float_rn(re[1]),
                           __double2float_rn(im[1])));
    }
    static __device__ __forceinline__ void load1(const void *p, size_t i, double &re, double &im)
    {
        const float2 x = __ldcs(reinterpret_cast<const float2 *>(p) + i);
        re = x.x; im = x.y;
    }
    static __device__ __forceinline__ void store1(void *p, size_t i, double re, double im, bool)
    {
        __stcs(reinterpret_cast<float2 *>(p) + i, make_float2(__double2float_rn(re), __double2float_rn(im)));
    }
};
template <> struct VisIO<double> {  // complex128: one pol = one 128-bit access
    static __device__ __forceinline__ void load1(const void *p, size_t i, double &re, double &im)
    {
        const double2 x = __ldcs(reinterpret_cast<const double2 *>(p) + i);
        re = x.x; im = x.y;
    }
    static __device__ __forceinline__ void store1(void *p, size_t i, double re, double im, bool single)
    {
        if (single) {   // (vis.astype(complex64)).astype(complex128), direction_rotate.py:244-245
            re = (double)__double2float_rn(re);
            im = (double)__double2float_rn(im);
        }
        __stcs(reinterpret_cast<double2 *>(p) + i, make_double2(re, im));
    }
    static __device__ __forceinline__ void load2(const void *p, size_t i, double *re, double *im)
    {
        load1(p, 2 * i, re[0], im[0]);
        load1(p, 2 * i + 1, re[1], im[1]);
    }
    static __device__ __forceinline__ void store2(void *p, size_t i, const double *re, const double *im, bool single)
    {
        store1(p, 2 * i, re[0], im[0], single);
        store1(p, 2 * i + 1, re[1], im[1], single);
    }
};

// numpy complex128 multiply: (a + ib)(c + is) = (ac - bs) + i(as + bc)
__device__ __forceinline__ void cmul(double a, double b, double c, double s, double &re, double &im)
{
    re = __dsub_rn(__dmul_rn(a, c), __dmul_rn(b, s));
    im = __dadd_rn(__dmul_rn(a, s), __dmul_rn(b, c));
}

// Phase of one (row, channel): the reference's ((2 pi d) f) (1/c), then sin/cos in fp64.
__device__ __forceinline__ void phasor_of(double two_pi_d, double f, double &c, double &s)
{
    const double y = __dmul_rn(__dmul_rn(two_pi_d, f), 1.0 / kSpeedOfLight);
    sincos(y, &s, &c);
}

// complex64 path: the phasor is rounded to fp32 and the product formed in fp32 (the fp32 bar is 1e-5 relative;
// the phase itself stays fp64).  One 128-bit access = 2 pols.
__device__ __forceinline__ float2 cmulf(float2 v, float c, float s)
{
    return make_float2(fmaf(v.x, c, -v.y * s), fmaf(v.x, s, v.y * c));
}

// One warp per (t, b) row: the lanes redundantly rotate the row's uvw and form 2 pi d (23 fp64 operations per row,
// not per sample), then walk the channels 32 at a time, CH chunks in flight, so every access is a contiguous
// 512-byte (complex64 x 2 pol) run.  NP = compile-time pol count (1, 2) or 0 = run-time n_pol.
template <typename T, int NP, int CH>
__global__ void __launch_bounds__(256)
dr_phasor_kernel(const void *vis, void *vis_rot, const double *__restrict__ uvw, const int *__restrict__ idx,
                 const double *__restrict__ freq, const double *__restrict__ rotmat,
                 const double *__restrict__ phase_rot, int n_time, int n_baseline, int n_chan, int n_pol,
                 unsigned n_rows, int end_slice, bool single)
{
    const unsigned r = blockIdx.x * 8u + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= n_rows) return;
    const int t = (int)(r / (unsigned)n_baseline);
    const int f_rot = idx[t], f_ph = idx[n_time + t];
    double tpd = nan("");
    if (f_rot >= 0 && f_ph >= 0) {
        double u, v, w;
        rotate_uvw(uvw + 3 * (size_t)r, rotmat + 9 * f_rot, u, v, w);
        const double *P = phase_rot + 3 * f_ph;
        double d = __dadd_rn(__dmul_rn(u, P[0]), __dmul_rn(v, P[1]));
        if (end_slice == 3) d = __dadd_rn(d, __dmul_rn(w, P[2]));
        tpd = __dmul_rn(6.283185307179586, d);
    }
    const size_t row0 = (size_t)r * n_chan;
    for (int c0 = 0; c0 < n_chan; c0 += 32 * CH) {
        if (NP == 2 && sizeof(T) == 4) {
            float4 x[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int c = c0 + 32 * j + lane;
                if (c < n_chan) x[j] = __ldcs(reinterpret_cast<const float4 *>(vis) + row0 + c);
            }
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int c = c0 + 32 * j + lane;
                if (c < n_chan) {
                    double cd, sd;
                    phasor_of(tpd, freq[c], cd, sd);
                    const float cf = __double2float_rn(cd), sf = __double2float_rn(sd);
                    const float2 a = cmulf(make_float2(x[j].x, x[j].y), cf, sf);
                    const float2 b = cmulf(make_float2(x[j].z, x[j].w), cf, sf);
                    __stcs(reinterpret_cast<float4 *>(vis_rot) + row0 + c, make_float4(a.x, a.y, b.x, b.y));
                }
            }
        } else if (NP > 0) {
            // compile-time pol count: all loads of the CH chunks are issued before the first sincos
            constexpr int NPC = NP > 0 ? NP : 1;
            using CT = typename Cplx<T>::type;
            CT x[CH][NPC];
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int c = c0 + 32 * j + lane;
#pragma unroll
                for (int p = 0; p < NPC; ++p)
                    if (c < n_chan) x[j][p] = __ldcs(reinterpret_cast<const CT *>(vis) + (row0 + c) * NPC + p);
            }
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int c = c0 + 32 * j + lane;
                if (c < n_chan) {
                    double cd, sd;
                    phasor_of(tpd, freq[c], cd, sd);
#pragma unroll
                    for (int p = 0; p < NPC; ++p) {
                        double ore, oim;
                        if (sizeof(T) == 4) {
                            const float2 o = cmulf(make_float2((float)x[j][p].x, (float)x[j][p].y),
                                                   __double2float_rn(cd), __double2float_rn(sd));
                            ore = o.x, oim = o.y;
                        } else {
                            cmul((double)x[j][p].x, (double)x[j][p].y, cd, sd, ore, oim);
                        }
                        VisIO<T>::store1(vis_rot, (row0 + c) * NPC + p, ore, oim, single);
                    }
                }
            }
        } else {
            for (int j = 0; j < CH; ++j) {
                const int c = c0 + 32 * j + lane;
                if (c < n_chan) {
                    double cd, sd;
                    phasor_of(tpd, freq[c], cd, sd);
                    for (int p = 0; p < n_pol; ++p) {
                        double re, im, ore, oim;
                        VisIO<T>::load1(vis, (row0 + c) * n_pol + p, re, im);
                        if (sizeof(T) == 4) {
                            const float2 o = cmulf(make_float2((float)re, (float)im), __double2float_rn(cd),
                                                   __double2float_rn(sd));
                            ore = o.x, oim = o.y;
                        } else {
                            cmul(re, im, cd, sd, ore, oim);
                        }
                        VisIO<T>::store1(vis_rot, (row0 + c) * n_pol + p, ore, oim, single);
                    }
                }
            }
        }
    }
}

template <typename T>
static int launch_phasor(const cngi_direction_rotate_args *a, const int *idx, cudaStream_t st)
{
    const long long n_rows = (long long)a->n_time * a->n_baseline;
    if (n_rows == 0 || a->n_chan == 0) return CNGI_OK;
    const unsigned blocks = (unsigned)ceil_div(n_rows, 8);
    const int end_slice = a->common_tangent_reprojection ? 2 : 3;
    const bool single = a->single_precision != 0;
#define CNGI_DR_LAUNCH(NP, CH)                                                                                     \
    dr_phasor_kernel<T, NP, CH><<<blocks, 256, 0, st>>>(a->vis, a->vis_rot, a->uvw, idx, a->freq_chan,              \
                                                        a->uvw_rotmat, a->phase_rotation, (int)a->n_time,           \
                                                        (int)a->n_baseline, (int)a->n_chan, (int)a->n_pol,          \
                                                        (unsigned)n_rows, end_slice, single)
    if (a->n_pol == 2) CNGI_DR_LAUNCH(2, 4);
    else if (a->n_pol == 1) CNGI_DR_LAUNCH(1, 4);
    else CNGI_DR_LAUNCH(0, 2);
#undef CNGI_DR_LAUNCH
    CNGI_CUDA_TRY(cudaGetLastError());
    return CNGI_OK;
}

}  // namespace cngi

extern "C" int cngi_b200_direction_rotate(const cngi_direction_rotate_args *a, void *stream)
{
    using namespace cngi;
    cudaStream_t st = (cudaStream_t)stream;
    CNGI_REQUIRE(a != nullptr, "direction_rotate: args is NULL");
    CNGI_REQUIRE(a->n_time >= 0 && a->n_baseline >= 0 && a->n_chan >= 0 && a->n_pol >= 1, "direction_rotate: bad shape");
    CNGI_REQUIRE(a->n_time * a->n_baseline < (1LL << 31) && a->n_chan < (1LL << 31), "direction_rotate: axis too long");
    if (a->n_time == 0 || a->n_baseline == 0) return CNGI_OK;   // empty chunk: nothing to do (pointers may be NULL)
    CNGI_REQUIRE(a->uvw && a->field && a->uvw_rotmat && a->phase_rotation && a->rot_field_id && a->n_field >= 1,
                 "direction_rotate: uvw, field, uvw_rotmat, phase_rotation, rot_field_id are required");
    CNGI_REQUIRE((a->vis == nullptr) == (a->vis_rot == nullptr), "direction_rotate: vis and vis_rot go together");
    CNGI_REQUIRE(a->vis == nullptr || a->freq_chan != nullptr, "direction_rotate: freq_chan is required with vis");
    CNGI_REQUIRE(a->precision == CNGI_F32 || a->precision == CNGI_F64, "direction_rotate: bad precision");

    if (int rc = tune_pool_once()) return rc;
    int *idx = nullptr;
    CNGI_CUDA_TRY(cudaMallocAsync((void **)&idx, (size_t)2 * a->n_time * sizeof(int), st));
    dr_field_index_kernel<<<(unsigned)ceil_div(a->n_time, 4), 128, 0, st>>>(a->field, (int)a->n_time, (int)a->n_baseline,
                                                                           a->rot_field_id, (int)a->n_field, idx,
                                                                           a->status);
    int rc = CNGI_OK;
    if (a->vis) rc = a->precision == CNGI_F32 ? launch_phasor<float>(a, idx, st) : launch_phasor<double>(a, idx, st);
    if (rc == CNGI_OK && a->uvw_rot) {   // after the phasor pass, so uvw_rot may alias uvw
        const long long n_rows = (long long)a->n_time * a->n_baseline;
        dr_uvw_kernel<<<(unsigned)ceil_div(n_rows, 256), 256, 0, st>>>(a->uvw, a->uvw_rot, idx, (int)a->n_baseline, n_rows,
                                                                      a->uvw_rotmat);
    }
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(idx, st);
    if (rc != CNGI_OK) return rc;
    CNGI_CUDA_TRY(e);
    return CNGI_OK;
}
