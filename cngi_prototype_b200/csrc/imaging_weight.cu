// imaging_weight.cu -- A2/A3/A4 of SURVEY.md section 8: Briggs / uniform imaging weights.
//   A2  density grid      _standard_grid_jit with do_imaging_weight, support 1   (_standard_grid.py:306-369,
//                         called from make_imaging_weight.py:153-161)
//   A3  briggs factors    calculate_briggs_parms                                 (make_imaging_weight.py:198-213)
//   A4  weight degrid     _standard_imaging_weight_degrid_jit                    (_standard_grid.py:466-518)
// One or two cells per sample.  A2 walks each (baseline, chan) track through time and run-length accumulates while the
// (cell, conjugate cell) pair is unchanged, so a slowly moving baseline issues one pair of REDG.F64 per cell crossing
// rather than per sample.
// A2 and A4 exist in two generations: the general kernels (iw_grid_kernel, iw_degrid_kernel / iw_degrid_mlp_kernel: any
// pol map, ragged channel groups, 64-bit strides) and the product kernels for what the reference's wrappers always hand
// in (iw_grid_fast_kernel, iw_degrid_fast_kernel: identity pol map, 1 or 2 pols, whole channel groups, aligned rows, a
// density of fewer than 2^31 cells) -- the launchers pick; CNGI_IW_GRID_OLD=1 / CNGI_IW_DEGRID_MLP=1 force the general
// ones (tests/test_gpu_imaging_weight.py holds the generations to each other).  DESIGN.md section 4.2 has the measurements.
#include "common.cuh"
#include <type_traits>
#include <cstdlib>

namespace cngi {

static inline bool env_flag(const char *name)   // development switches (A/B timing of kernel generations)
{
    static_assert(true, "");
    const char *e = getenv(name);
    return e && e[0] && e[0] != '0';
}

struct IwParams {
    int n_time, n_baseline, n_chan, n_pol;
    int n_ic, n_ip, n_u, n_v;
    const void *weight;
    const double *uvw;
    const double *freq;
    const int64_t *chan_map;
    const int64_t *pol_map;
    double *density;
    double *sum_weight;
    double dl, dm;
    int chan_mode;
    int seg_len, n_seg;
    // degrid
    const double *bf;
    long long ds_u, ds_v, ds_c, ds_p;
    void *out;
    const double *scale;   // [2, n_chan] uv_scale table
    int n_pol_out;         // pol planes updated by iw_grid (1 when first_pol_only)
    int pol_shared;        // degrid: density planes and Briggs factors are identical across pol (one gather, one division)
};

__device__ __forceinline__ int iw_chan_of(const IwParams &p, int c)
{
    if (p.chan_mode == CNGI_CHAN_CUBE) return c;
    if (p.chan_mode == CNGI_CHAN_CONTINUUM) return 0;
    return (int)p.chan_map[c];
}

// thread <-> (time segment, baseline, group of G consecutive channels); group index fastest, so a warp reads
// 32*G consecutive channels of one row (coalesced).  The thread walks time (outer) and its G channels (inner,
// boustrophedon so consecutive samples stay neighbours in the uv plane) and run-length accumulates while the
// (plane, cell, conjugate cell) key is unchanged.  G > 1 is only used when all channels share ONE image plane
// (continuum): neighbouring channels of a baseline then fall into the same cell, which removes most same-address
// reductions.  NP = compile-time pol count (1 / 2 with the identity pol_map), 0 = generic.
template <typename T, int G, int NP> __global__ void __launch_bounds__(256) iw_grid_kernel(IwParams p)
{
    const int n_pol = NP ? NP : p.n_pol;
    const long long plane_cells = (long long)p.n_u * p.n_v;
    const int n_cg = (p.n_chan + G - 1) / G;
    const long long n_items = (long long)p.n_seg * p.n_baseline * n_cg;
    long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = item < n_items;   // no early return: the warp reduces sum_weight together at the end
    if (!in_range) item = 0;
    const int cg = (int)(item % n_cg);
    const long long r = item / n_cg;
    const int b = (int)(r % p.n_baseline);
    const int seg = (int)(r / p.n_baseline);
    const int t_lo = seg * p.seg_len, t_hi = in_range ? min(p.n_time, t_lo + p.seg_len) : t_lo;
    const int c0 = cg * G;
    const int ng = min(G, p.n_chan - c0);
    const int plane = (G > 1) ? 0 : iw_chan_of(p, c0);   // G > 1 <=> continuum
    double *const plane_base = p.density + (long long)plane * p.n_ip * plane_cells;
    const bool average = n_pol >= 2;               // (n_pol >= 2) and do_imaging_weight, :328-330
    const double mid_u = (double)(p.n_u / 2), mid_v = (double)(p.n_v / 2);
    const double *us = p.scale + c0, *vs = p.scale + p.n_chan + c0;   // uv_scale table (L1 resident)

    int cur_u = -1, cur_v = 0, cur_cu = 0, cur_cv = 0;
    double acc = 0.0, sw = 0.0;

    auto flush = [&]() {
        if (cur_u < 0 || acc == 0.0) return;
        const bool conj_ok = cur_cu >= 0 && cur_cu < p.n_u && cur_cv >= 0 && cur_cv < p.n_v;
        const int cell = cur_u * p.n_v + cur_v, ccell = cur_cu * p.n_v + cur_cv;   // n_u * n_v < 2^31 (checked by the launcher)
        if (NP) {
#pragma unroll
            for (int ip = 0; ip < (NP ? NP : 1); ++ip) {
                if (ip < p.n_pol_out) {
                    atomicAdd(plane_base + ip * plane_cells + cell, acc);
                    if (conj_ok) atomicAdd(plane_base + ip * plane_cells + ccell, acc);
                }
            }
        } else {
            for (int ip = 0; ip < p.n_pol_out; ++ip) {
                const int a_pol = p.pol_map ? (int)p.pol_map[ip] : ip;
                atomicAdd(plane_base + a_pol * plane_cells + cell, acc);
                if (conj_ok) atomicAdd(plane_base + a_pol * plane_cells + ccell, acc);
            }
        }
        acc = 0.0;
    };

    for (int t = t_lo; t < t_hi; ++t) {
        const long long tb = (long long)t * p.n_baseline + b;
        const double uu = p.uvw[tb * 3], vv = p.uvw[tb * 3 + 1];
        const T *wrow = (const T *)p.weight + (tb * p.n_chan + c0) * n_pol;
        double wd[G];
        if (NP == 2 && sizeof(T) == 4 && G >= 2 && ng == G && ((reinterpret_cast<uintptr_t>(wrow) & 15) == 0)) {
            // the thread's G channels x 2 pols are 8 G contiguous bytes: 128-bit loads (two channels each)
#pragma unroll
            for (int g = 0; g < G; g += 2) {
                const float4 w4 = *reinterpret_cast<const float4 *>(wrow + g * 2);
                wd[g] = __dmul_rn(__dadd_rn((double)w4.x, (double)w4.y), 0.5);
                wd[g + 1] = __dmul_rn(__dadd_rn((double)w4.z, (double)w4.w), 0.5);
            }
        } else
#pragma unroll
        for (int g = 0; g < G; ++g) {   // issue all loads of the row before the dependent math
            wd[g] = 0.0;
            if (g < ng) {
                if (average) {
                    double w0, w1;
                    if (n_pol == 2) {
                        if (sizeof(T) == 4) {
                            const float2 w2 = *reinterpret_cast<const float2 *>(wrow + g * 2);
                            w0 = (double)w2.x, w1 = (double)w2.y;
                        } else {
                            const double2 w2 = *reinterpret_cast<const double2 *>(wrow + g * 2);
                            w0 = w2.x, w1 = w2.y;
                        }
                    } else {
                        w0 = (double)wrow[g * n_pol], w1 = (double)wrow[g * n_pol + 1];
                    }
                    wd[g] = __dmul_rn(__dadd_rn(w0, w1), 0.5);   // == /2.0 exactly
                } else {
                    wd[g] = (double)wrow[g];
                }
            }
        }
        // g is a compile-time constant in both unrolled loops below (a run-time index would push wd[] to local memory)
        auto sample = [&](int g, double w) {
            if (g >= ng) return;
            CellPos cp;
            const double su = us[g], sv = vs[g];
            if (!locate_centre(uu, vv, su, sv, p.n_u, p.n_v, cp)) return;
            if (!stamp_inside(cp.uc, cp.vc, 0, p.n_u, p.n_v)) return;
            if (isnan(w) || w == 0.0) return;
            // conjugate cell: int(-u + centre + 0.5)   (:309-318)
            const double un = -__dmul_rn(uu, su), vn = -__dmul_rn(vv, sv);
            const int cu = __double2int_rz(__dadd_rn(__dadd_rn(un, mid_u), 0.5));
            const int cv = __double2int_rz(__dadd_rn(__dadd_rn(vn, mid_v), 0.5));
            if (cp.uc != cur_u || cp.vc != cur_v || cu != cur_cu || cv != cur_cv) {
                flush();
                cur_u = cp.uc, cur_v = cp.vc, cur_cu = cu, cur_cv = cv;
            }
            acc += w;
            sw += w + w;   // sum_weight gets sel_weight*norm twice (:366-369), norm == cgk_1D[0] == 1
        };
        if (t & 1) {   // boustrophedon
#pragma unroll
            for (int gi = 0; gi < G; ++gi) sample(G - 1 - gi, wd[G - 1 - gi]);
        } else {
#pragma unroll
            for (int gi = 0; gi < G; ++gi) sample(gi, wd[gi]);
        }
    }
    flush();
    // sum_weight: one reduction per plane per warp (all of a thread's channels share its plane)
    for (int ip = 0; ip < p.n_pol_out; ++ip) {   // uniform trip count, so the warp stays converged
        const int a_pol = (!NP && p.pol_map) ? (int)p.pol_map[ip] : ip;
        warp_grouped_add(p.sum_weight, plane * p.n_ip + a_pol, sw, in_range && sw != 0.0);
    }
}

// Product path of A2 (identity pol_map, 1 or 2 pols, whole channel groups, 16-byte aligned rows).  ncu on the kernel above
// (profiles/r01_imaging_weight_kernels.txt): 38 % issue utilisation, 65 % of the stall samples on the long scoreboard -- a
// time step's weights were loaded right before their use -- and ~85 warp instructions per sample, because every sample ran
// its own chain of early-outs with the flush inlined behind it.  Here the NEXT time step's weights and (u, v) are loaded
// before the current one is processed, all G cells of a step are located first in straight-line code (G independent fp64
// chains), and only then merged into the running (cell, conjugate cell) accumulator with integer compares.
#ifndef CNGI_IW_GRID_HALF
#define CNGI_IW_GRID_HALF 4
#endif
#ifndef CNGI_IW_GRID_MINB
#define CNGI_IW_GRID_MINB 2
#endif
template <typename T> struct IwPair;
template <> struct IwPair<float> { using type = float2; };
template <> struct IwPair<double> { using type = double2; };

template <typename T, int G, int NP> __global__ void __launch_bounds__(256, CNGI_IW_GRID_MINB) iw_grid_fast_kernel(IwParams p)
{
    using V2 = typename IwPair<T>::type;
    const long long plane_cells = (long long)p.n_u * p.n_v;
    const int n_cg = p.n_chan / G;
    const long long n_items = (long long)p.n_seg * p.n_baseline * n_cg;
    long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = item < n_items;   // no early return: the warp reduces sum_weight together at the end
    if (!in_range) item = 0;
    const int cg = (int)(item % n_cg);
    const long long r = item / n_cg;
    const int b = (int)(r % p.n_baseline);
    const int seg = (int)(r / p.n_baseline);
    const int t_lo = seg * p.seg_len, t_hi = in_range ? min(p.n_time, t_lo + p.seg_len) : t_lo;
    const int c0 = cg * G;
    const int plane = (G > 1) ? 0 : iw_chan_of(p, c0);   // G > 1 <=> continuum
    double *const plane_base = p.density + (long long)plane * p.n_ip * plane_cells;
    const double mid_u = (double)(p.n_u / 2), mid_v = (double)(p.n_v / 2);
    // uv_scale table, transposed in shared memory to [u | v][g][cg]: the lanes of a warp (consecutive cg) read consecutive
    // doubles.  Straight from the global table the 16 loads of a time step were strided by G doubles across the lanes --
    // 8 cache lines per request, 28 % of the stall samples of the first version of this kernel (ncu).
    extern __shared__ double iw_scale_sm[];
    for (int i = threadIdx.x; i < 2 * p.n_chan; i += blockDim.x) {
        const int uv = i / p.n_chan, c = i - uv * p.n_chan;
        iw_scale_sm[uv * p.n_chan + (c % G) * n_cg + c / G] = p.scale[i];
    }
    __syncthreads();
    const double *us = iw_scale_sm + cg, *vs = iw_scale_sm + p.n_chan + cg;   // channel c0 + g at [g * n_cg]

    const long long row_stride = (long long)p.n_baseline * p.n_chan * NP;
    const T *wrow = (const T *)p.weight + (((long long)t_lo * p.n_baseline + b) * p.n_chan + c0) * NP;
    const double *uvp = p.uvw + ((long long)t_lo * p.n_baseline + b) * 3;
    const long long uv_stride = (long long)p.n_baseline * 3;

    // Running (cell, conjugate cell, sum) of the lane.  (Measured and dropped: holding a retired triple back until some lane of
    // the warp needs the slot again, so that more lanes share a REDG -- 2.5 M -> 1.0 M reduction instructions, same time:
    // the kernel is not bound by the reductions; plane 0 + plane 1 instead of plane 0 alone costs 0.03 ms.)
    int cur = -1, cur_c = -1;     // (-1: none / off the grid)
    double acc = 0.0, sw = 0.0;
    const bool both_planes = NP == 2 && p.n_pol_out == 2;
    auto reduce_out = [&](int cell, int ccell, double v) {   // only reads: the state changes around it are selects
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) {
            if (ip == 0 || both_planes) {
                atomicAdd(plane_base + ip * plane_cells + cell, v);
                if (ccell >= 0) atomicAdd(plane_base + ip * plane_cells + ccell, v);
            }
        }
    };

    T raw[G][NP];          // the NEXT time step's weights
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) raw[g][ip] = (T)0;
    double nu = 0.0, nv = 0.0;
    auto load_row = [&]() {
        nu = uvp[0], nv = uvp[1];
        if (NP == 2 && sizeof(T) == 4 && G >= 2) {   // two channels x 2 pols per 128-bit load
#pragma unroll
            for (int g = 0; g < G; g += 2) {
                const float4 w4 = *reinterpret_cast<const float4 *>(wrow + g * 2);
                raw[g][0] = (T)w4.x, raw[g][NP - 1] = (T)w4.y, raw[g + 1 < G ? g + 1 : g][0] = (T)w4.z,
                raw[g + 1 < G ? g + 1 : g][NP - 1] = (T)w4.w;
            }
        } else if (NP == 2) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const V2 w2 = *reinterpret_cast<const V2 *>(wrow + g * 2);
                raw[g][0] = w2.x, raw[g][NP - 1] = w2.y;
            }
        } else {
#pragma unroll
            for (int g = 0; g < G; ++g) raw[g][0] = wrow[g];
        }
        wrow += row_stride;
        uvp += uv_stride;
    };

    if (t_lo < t_hi) load_row();
    for (int t = t_lo; t < t_hi; ++t) {
        T w_now[G][NP];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int ip = 0; ip < NP; ++ip) w_now[g][ip] = raw[g][ip];
        const double uu = nu, vv = nv;
        if (t + 1 < t_hi) load_row();   // in flight while this step is processed
        // ---- H cells at a time, branch-free (locate_centre + stamp_inside + the conjugate cell, :299-318), then merged into
        //      the running accumulator; boustrophedon so that consecutive samples stay uv neighbours ----
        constexpr int H = G < CNGI_IW_GRID_HALF ? G : CNGI_IW_GRID_HALF;
        const bool back = (t - t_lo) & 1;
#pragma unroll
        for (int hh = 0; hh < G / H; ++hh) {
            int cell[H], ccell[H];
            double wv[H];
            auto locate = [&](int k, int g) {   // k, g compile-time constants after unrolling
                const double w = NP == 2 ? __dmul_rn(__dadd_rn((double)w_now[g][0], (double)w_now[g][NP - 1]), 0.5)   // == /2.0 exactly
                                         : (double)w_now[g][0];
                const double u = __dmul_rn(uu, us[g * n_cg]), v = __dmul_rn(vv, vs[g * n_cg]);
                const int uc = __double2int_rz(__dadd_rn(__dadd_rn(u, mid_u), 0.5));
                const int vc = __double2int_rz(__dadd_rn(__dadd_rn(v, mid_v), 0.5));
                const int cu = __double2int_rz(__dadd_rn(__dadd_rn(-u, mid_u), 0.5));   // int(-u + centre + 0.5); -(u) is exact
                const int cv = __double2int_rz(__dadd_rn(__dadd_rn(-v, mid_v), 0.5));
                const bool ok = (u == u) && (v == v) && (uc < p.n_u) && (vc < p.n_v) && (uc >= 0) && (vc >= 0) && (w == w) &&
                                (w != 0.0);
                const bool cok = (cu >= 0) && (cu < p.n_u) && (cv >= 0) && (cv < p.n_v);
                cell[k] = ok ? uc * p.n_v + vc : -1;          // n_u * n_v < 2^31 (checked by the launcher)
                ccell[k] = cok ? cu * p.n_v + cv : -1;
                wv[k] = ok ? w : 0.0;
            };
            auto merge = [&](int k) {
                const bool moved = cell[k] >= 0 && (cell[k] != cur || ccell[k] != cur_c);
                const bool retire = moved && acc != 0.0;   // (cur >= 0 whenever acc != 0)
                if (retire) reduce_out(cur, cur_c, acc);
                acc = moved ? 0.0 : acc;
                cur = moved ? cell[k] : cur;
                cur_c = moved ? ccell[k] : cur_c;
                acc += wv[k];
                sw += wv[k];
            };
            if (back) {   // the halves are walked downwards on odd steps
#pragma unroll
                for (int k = 0; k < H; ++k) locate(k, (G / H - 1 - hh) * H + k);
            } else {
#pragma unroll
                for (int k = 0; k < H; ++k) locate(k, hh * H + k);
            }
            if (back) {
#pragma unroll
                for (int k = H - 1; k >= 0; --k) merge(k);
            } else {
#pragma unroll
                for (int k = 0; k < H; ++k) merge(k);
            }
        }
    }
    if (acc != 0.0) reduce_out(cur, cur_c, acc);
    // sum_weight gets sel_weight * norm twice (:366-369), norm == cgk_1D[0] == 1; one reduction per plane per warp
    sw += sw;
    for (int ip = 0; ip < p.n_pol_out; ++ip)   // uniform trip count, so the warp stays converged
        warp_grouped_add(p.sum_weight, plane * p.n_ip + ip, sw, in_range && sw != 0.0);
}

// sum of squares per plane -> bf[0][plane] (used as the accumulator, finalised below)
__global__ void __launch_bounds__(256) iw_sumsq_kernel(const double *density, double *acc, long long n_cells)
{
    const int plane = blockIdx.y;
    const double *d = density + (long long)plane * n_cells;
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += (long long)gridDim.x * blockDim.x) {
        const double x = d[i];
        s = fma(x, x, s);
    }
    __shared__ double part[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < 8 ? part[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) atomicAdd(acc + plane, s);
    }
}

__global__ void iw_briggs_finalize_kernel(double *bf, const double *sum_weight, long long n_planes, double k2, int weighting)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_planes) return;
    if (weighting == 0) {   // briggs: (5*10^-robust)^2 / (sum(rho^2)/sum_weight), f1 = 1
        bf[i] = __ddiv_rn(k2, __ddiv_rn(bf[i], sum_weight[i]));
        bf[n_planes + i] = 1.0;
    } else {                // uniform: f0 = 1, f1 = 0
        bf[i] = 1.0;
        bf[n_planes + i] = 0.0;
    }
}

// block = (CX channels) x (256 / CX rows); thread <-> one (row, chan) sample, all pols.  No integer division, the uvw
// row is a warp-wide broadcast load, weights and outputs are coalesced along the channel axis.
template <typename T, int NP> __global__ void __launch_bounds__(256) iw_degrid_kernel(IwParams p)
{
    const int n_pol = NP ? NP : p.n_pol;
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    const long long tb = (long long)blockIdx.x * blockDim.y + threadIdx.y;
    if (c >= p.n_chan || tb >= (long long)p.n_time * p.n_baseline) return;
    const long long idx = tb * p.n_chan + c;
    T *out = (T *)p.out + idx * n_pol;
    const T *nat = (const T *)p.weight + idx * n_pol;
    T natv[NP ? NP : 1];
    if (NP == 2) {   // one 8/16-byte load for the pol pair
        if (sizeof(T) == 4) {
            const float2 w2 = *reinterpret_cast<const float2 *>(nat);
            natv[0] = (T)w2.x, natv[NP - 1] = (T)w2.y;
        } else {
            const double2 w2 = *reinterpret_cast<const double2 *>(nat);
            natv[0] = (T)w2.x, natv[NP - 1] = (T)w2.y;
        }
    } else if (NP == 1) {
        natv[0] = nat[0];
    }
    CellPos cp;
    bool ok = locate_centre(p.uvw[tb * 3], p.uvw[tb * 3 + 1], p.scale[c], p.scale[p.n_chan + c], p.n_u, p.n_v, cp);
    if (ok) ok = stamp_inside(cp.uc, cp.vc, 0, p.n_u, p.n_v);
    T res[NP ? NP : 1];
    if (!ok) {   // off-grid or NaN uv: output stays 0 (:460,493,502)
        if (NP) {
#pragma unroll
            for (int ip = 0; ip < (NP ? NP : 1); ++ip) res[ip] = (T)0;
        } else {
            for (int ip = 0; ip < n_pol; ++ip) out[ip] = (T)0;
            return;
        }
    } else {
        const int a_chan = iw_chan_of(p, c);
        const long long cell = cp.uc * p.ds_u + cp.vc * p.ds_v + a_chan * p.ds_c;
        if (NP) {
            const double avg = NP == 2 ? __dmul_rn(__dadd_rn((double)natv[0], (double)natv[NP - 1]), 0.5) : 0.0;   // == /2.0
#pragma unroll
            for (int ip = 0; ip < (NP ? NP : 1); ++ip) {
                double iw = NP == 2 ? avg : (double)natv[ip];   // :508-511
                const double w = (double)natv[ip];
                if (!isnan(w) && w != 0.0) {
                    const double rho = p.density[cell + ip * p.ds_p];
                    if (!isnan(rho) && rho != 0.0) {
                        const double den = __dadd_rn(__dmul_rn(p.bf[a_chan * p.n_ip + ip], rho),
                                                     p.bf[((long long)p.n_ic + a_chan) * p.n_ip + ip]);   // :515-516
                        iw = sizeof(T) == 4 ? (double)__fdiv_rn((float)iw, (float)den) : __ddiv_rn(iw, den);
                    }
                }
                res[ip] = (T)iw;
            }
        } else {
            const double avg = n_pol == 2 ? __dmul_rn(__dadd_rn((double)nat[0], (double)nat[1]), 0.5) : 0.0;
            for (int ip = 0; ip < n_pol; ++ip) {
                const int a_pol = p.pol_map ? (int)p.pol_map[ip] : ip;
                double iw = n_pol == 2 ? avg : (double)nat[ip];
                const double w = (double)nat[ip];
                if (!isnan(w) && w != 0.0) {
                    const double rho = p.density[cell + a_pol * p.ds_p];
                    if (!isnan(rho) && rho != 0.0) {
                        const double den = __dadd_rn(__dmul_rn(p.bf[a_chan * p.n_ip + a_pol], rho),
                                                     p.bf[((long long)p.n_ic + a_chan) * p.n_ip + a_pol]);
                        iw = sizeof(T) == 4 ? (double)__fdiv_rn((float)iw, (float)den) : __ddiv_rn(iw, den);
                    }
                }
                out[ip] = (T)iw;
            }
            return;
        }
    }
    if (NP == 2) {
        if (sizeof(T) == 4)
            *reinterpret_cast<float2 *>(out) = make_float2((float)res[0], (float)res[NP - 1]);
        else
            *reinterpret_cast<double2 *>(out) = make_double2((double)res[0], (double)res[NP - 1]);
    } else if (NP == 1) {
        out[0] = res[0];
    }
}

// Fast path of A4 for the identity pol_map with 1 or 2 pols: R samples (rows tb, tb + stride, ...) per thread, written as
// three passes so that all R uvw / weight loads, then all R density gathers, are in flight together -- the plain kernel
// above is bound by its dependent uvw -> cell -> gather latency chain (63 % long-scoreboard stalls).
template <typename T, int NP, int R> __global__ void __launch_bounds__(256) iw_degrid_mlp_kernel(IwParams p)
{
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    const long long rows = (long long)p.n_time * p.n_baseline;
    const long long row0 = ((long long)blockIdx.x * R) * blockDim.y + threadIdx.y;   // rows row0 + k * blockDim.y
    if (c >= p.n_chan) return;
    const double us = p.scale[c], vs = p.scale[p.n_chan + c];
    const int a_chan = iw_chan_of(p, c);
    double f0[NP], f1[NP];
#pragma unroll
    for (int ip = 0; ip < NP; ++ip) {
        f0[ip] = p.bf[a_chan * p.n_ip + ip];
        f1[ip] = p.bf[((long long)p.n_ic + a_chan) * p.n_ip + ip];
    }
    double uu[R], vv[R];
    T nat[R][NP];
    bool live[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {   // pass 1: independent loads
        const long long tb = row0 + (long long)k * blockDim.y;
        live[k] = tb < rows;
        uu[k] = vv[k] = 0.0;
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) nat[k][ip] = (T)0;
        if (live[k]) {
            uu[k] = p.uvw[tb * 3];
            vv[k] = p.uvw[tb * 3 + 1];
            const T *np_ = (const T *)p.weight + (tb * p.n_chan + c) * NP;
            if (NP == 2) {
                if (sizeof(T) == 4) {
                    const float2 w2 = *reinterpret_cast<const float2 *>(np_);
                    nat[k][0] = (T)w2.x, nat[k][NP - 1] = (T)w2.y;
                } else {
                    const double2 w2 = *reinterpret_cast<const double2 *>(np_);
                    nat[k][0] = (T)w2.x, nat[k][NP - 1] = (T)w2.y;
                }
            } else {
                nat[k][0] = np_[0];
            }
        }
    }
    double rho[R][NP];
    bool ok[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {   // pass 2: cells, then the gathers
        CellPos cp;
        ok[k] = live[k] && locate_centre(uu[k], vv[k], us, vs, p.n_u, p.n_v, cp);
        if (ok[k]) ok[k] = stamp_inside(cp.uc, cp.vc, 0, p.n_u, p.n_v);
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) rho[k][ip] = 0.0;
        if (ok[k]) {
            const long long cell = cp.uc * p.ds_u + cp.vc * p.ds_v + a_chan * p.ds_c;
            if (NP == 2 && p.pol_shared) {
                rho[k][0] = p.density[cell];   // every pol sees this value
            } else {
#pragma unroll
                for (int ip = 0; ip < NP; ++ip) {
                    const double w = (double)nat[k][ip];
                    if (!isnan(w) && w != 0.0) rho[k][ip] = p.density[cell + ip * p.ds_p];   // rho is only used in that case
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {   // pass 3: Briggs division, stores
        if (!live[k]) continue;
        const long long tb = row0 + (long long)k * blockDim.y;
        T res[NP];
        const double avg = NP == 2 ? __dmul_rn(__dadd_rn((double)nat[k][0], (double)nat[k][NP - 1]), 0.5) : 0.0;   // == /2.0
        if (NP == 2 && p.pol_shared) {   // same value and factors for both pols: one division (bit-identical results)
            double q = avg;
            const double r = rho[k][0];
            if (ok[k] && !isnan(r) && r != 0.0) {
                const double den = __dadd_rn(__dmul_rn(f0[0], r), f1[0]);
                q = sizeof(T) == 4 ? (double)__fdiv_rn((float)avg, (float)den) : __ddiv_rn(avg, den);
            }
#pragma unroll
            for (int ip = 0; ip < NP; ++ip) {
                const double w = (double)nat[k][ip];
                res[ip] = (T)(ok[k] ? ((!isnan(w) && w != 0.0) ? q : avg) : 0.0);
            }
        } else
#pragma unroll
        for (int ip = 0; ip < NP; ++ip) {
            double iw = 0.0;   // off-grid or NaN uv: 0 (:460,493,502)
            if (ok[k]) {
                iw = NP == 2 ? avg : (double)nat[k][ip];   // :508-511
                const double w = (double)nat[k][ip], r = rho[k][ip];
                if (!isnan(w) && w != 0.0 && !isnan(r) && r != 0.0) {
                    const double den = __dadd_rn(__dmul_rn(f0[ip], r), f1[ip]);   // :515-516
                    iw = sizeof(T) == 4 ? (double)__fdiv_rn((float)iw, (float)den) : __ddiv_rn(iw, den);
                }
            }
            res[ip] = (T)iw;
        }
        T *out = (T *)p.out + (tb * p.n_chan + c) * NP;
        if (NP == 2) {
            if (sizeof(T) == 4)
                *reinterpret_cast<float2 *>(out) = make_float2((float)res[0], (float)res[NP - 1]);
            else
                *reinterpret_cast<double2 *>(out) = make_double2((double)res[0], (double)res[NP - 1]);
        } else {
            out[0] = res[0];
        }
    }
}


// Product path of A4 (identity pol_map, 1 or 2 pols, density smaller than 2^31 elements): the mlp kernel above retired 180 warp
// instructions per sample at 72 % issue utilisation (ncu, profiles/r01_imaging_weight_kernels.txt) -- 64-bit stride
// arithmetic for every gather, per-sample range predicates on every address, a branch around every step.  Here the cell
// index is 32-bit, addresses advance by a constant stride, range checks exist only in the launch's last row block, and the
// three passes (loads / cells + gathers / Briggs division + stores) are straight-line code with selects: a sample that is off
// the grid gathers cell 0 and discards it.  Results are bit-identical to iw_degrid_kernel's.
template <typename T, int NP, int R, bool SHARED> __global__ void __launch_bounds__(256) iw_degrid_fast_kernel(IwParams p)
{
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= p.n_chan) return;
    const long long rows = (long long)p.n_time * p.n_baseline;
    const long long row0 = ((long long)blockIdx.x * R) * blockDim.y + threadIdx.y;   // rows row0 + k * blockDim.y
    const double us = p.scale[c], vs = p.scale[p.n_chan + c];
    const int a_chan = iw_chan_of(p, c);
    double f0[NP], f1[NP];
#pragma unroll
    for (int ip = 0; ip < NP; ++ip) {
        f0[ip] = p.bf[a_chan * p.n_ip + ip];
        f1[ip] = p.bf[((long long)p.n_ic + a_chan) * p.n_ip + ip];
    }
    const double mid_u = (double)(p.n_u / 2), mid_v = (double)(p.n_v / 2);
    const int ds_u = (int)p.ds_u, ds_v = (int)p.ds_v, ds_p = (int)p.ds_p;
    const double *dens = p.density + (long long)a_chan * p.ds_c;
    asm volatile("" : "+l"(dens));   // keep the channel offset folded into the base: ptxas otherwise redoes the 64-bit sum per gather
    const long long e_stride = (long long)blockDim.y * p.n_chan * NP;
    const T *const wp = (const T *)p.weight + (row0 * p.n_chan + c) * NP;
    T *const op = (T *)p.out + (row0 * p.n_chan + c) * NP;
    const double *const uvp = p.uvw + row0 * 3;
    const int uv_stride = (int)blockDim.y * 3;

    auto body = [&](auto checked) {
        constexpr bool CHK = decltype(checked)::value;   // only the last row block of the launch can run past `rows`
        double uu[R], vv[R];
        T nat[R][NP];
#pragma unroll
        for (int k = 0; k < R; ++k) {   // pass 1: independent loads
            const bool live = !CHK || (row0 + (long long)k * blockDim.y < rows);
            uu[k] = vv[k] = 0.0;
#pragma unroll
            for (int ip = 0; ip < NP; ++ip) nat[k][ip] = (T)0;
            if (live) {
                uu[k] = uvp[k * uv_stride];
                vv[k] = uvp[k * uv_stride + 1];
                const T *np_ = wp + k * e_stride;
                if (NP == 2) {
                    if (sizeof(T) == 4) {
                        const float2 w2 = *reinterpret_cast<const float2 *>(np_);
                        nat[k][0] = (T)w2.x, nat[k][NP - 1] = (T)w2.y;
                    } else {
                        const double2 w2 = *reinterpret_cast<const double2 *>(np_);
                        nat[k][0] = (T)w2.x, nat[k][NP - 1] = (T)w2.y;
                    }
                } else {
                    nat[k][0] = np_[0];
                }
            }
        }
        double rho[R][SHARED ? 1 : NP];
        bool ok[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {   // pass 2: cells (locate_centre + stamp_inside, branch-free), then the gathers
            const double u = __dmul_rn(uu[k], us), v = __dmul_rn(vv[k], vs);
            const int uc = __double2int_rz(__dadd_rn(__dadd_rn(u, mid_u), 0.5));
            const int vc = __double2int_rz(__dadd_rn(__dadd_rn(v, mid_v), 0.5));
            ok[k] = (u == u) && (v == v) && (uc < p.n_u) && (vc < p.n_v) && (uc >= 0) && (vc >= 0);
            const int cell = ok[k] ? uc * ds_u + vc * ds_v : 0;
#pragma unroll
            for (int ip = 0; ip < (SHARED ? 1 : NP); ++ip) rho[k][ip] = dens[cell + ip * ds_p];
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {   // pass 3: Briggs division, stores
            if (CHK && !(row0 + (long long)k * blockDim.y < rows)) continue;
            T res[NP];
            const double avg = NP == 2 ? __dmul_rn(__dadd_rn((double)nat[k][0], (double)nat[k][NP - 1]), 0.5) : 0.0;   // == /2.0
            if (SHARED) {   // same density value and factors for both pols: one division
                const double r = rho[k][0];
                const double den = __dadd_rn(__dmul_rn(f0[0], r), f1[0]);   // :515-516
                const double quo = sizeof(T) == 4 ? (double)__fdiv_rn((float)avg, (float)den) : __ddiv_rn(avg, den);
                const double q = (ok[k] && r == r && r != 0.0) ? quo : avg;
#pragma unroll
                for (int ip = 0; ip < NP; ++ip) {
                    const double w = (double)nat[k][ip];
                    res[ip] = (T)(ok[k] ? ((w == w && w != 0.0) ? q : avg) : 0.0);
                }
            } else {
#pragma unroll
                for (int ip = 0; ip < NP; ++ip) {
                    const double w = (double)nat[k][ip], r = rho[k][SHARED ? 0 : ip];
                    const double num = NP == 2 ? avg : w;   // :508-511
                    const double den = __dadd_rn(__dmul_rn(f0[ip], r), f1[ip]);
                    const double quo = sizeof(T) == 4 ? (double)__fdiv_rn((float)num, (float)den) : __ddiv_rn(num, den);
                    const bool divide = (w == w) && w != 0.0 && (r == r) && r != 0.0;
                    res[ip] = (T)(ok[k] ? (divide ? quo : num) : 0.0);   // off-grid or NaN uv: 0 (:460,493,502)
                }
            }
            T *out = op + k * e_stride;
            if (NP == 2) {
                if (sizeof(T) == 4)
                    *reinterpret_cast<float2 *>(out) = make_float2((float)res[0], (float)res[NP - 1]);
                else
                    *reinterpret_cast<double2 *>(out) = make_double2((double)res[0], (double)res[NP - 1]);
            } else {
                out[0] = res[0];
            }
        }
    };
    if (row0 + (long long)(R - 1) * blockDim.y < rows)
        body(std::false_type{});
    else
        body(std::true_type{});
}

}  // namespace cngi

extern "C" int cngi_b200_imaging_weight_grid(const cngi_iw_grid_args *a, void *stream)
{
    using namespace cngi;
    CNGI_REQUIRE(a != nullptr, "imaging_weight_grid: null args");
    CNGI_REQUIRE(a->weight && a->uvw && a->freq_chan && a->density && a->sum_weight, "imaging_weight_grid: null array pointer");
    CNGI_REQUIRE(a->precision == CNGI_F32 || a->precision == CNGI_F64, "imaging_weight_grid: bad precision");
    CNGI_REQUIRE(a->n_u > 0 && a->n_v > 0 && a->n_u < (1 << 24) && a->n_v < (1 << 24), "imaging_weight_grid: bad grid size");
    CNGI_REQUIRE(a->chan_mode != CNGI_CHAN_GENERAL || a->chan_map, "imaging_weight_grid: chan_map is null");
    CNGI_REQUIRE(a->n_time * a->n_baseline < (1LL << 31), "imaging_weight_grid: too many rows");
    CNGI_REQUIRE(a->n_u * a->n_v < (1LL << 31), "imaging_weight_grid: n_u*n_v overflows int32");
    if (a->n_time == 0 || a->n_baseline == 0 || a->n_chan == 0 || a->n_pol == 0) return CNGI_OK;
    IwParams p{};
    p.n_time = (int)a->n_time, p.n_baseline = (int)a->n_baseline, p.n_chan = (int)a->n_chan, p.n_pol = (int)a->n_pol;
    p.n_ic = (int)a->n_imag_chan, p.n_ip = (int)a->n_imag_pol, p.n_u = (int)a->n_u, p.n_v = (int)a->n_v;
    p.weight = a->weight, p.uvw = a->uvw, p.freq = a->freq_chan, p.chan_map = a->chan_map, p.pol_map = a->pol_map;
    p.density = a->density, p.sum_weight = a->sum_weight, p.dl = a->delta_lm[0], p.dm = a->delta_lm[1];
    p.chan_mode = a->chan_mode;
    p.n_pol_out = (a->first_pol_only && p.n_pol >= 2) ? 1 : p.n_pol;
    // channels per thread: only useful when neighbouring channels share a plane
    int G = 1;
    if (p.chan_mode == CNGI_CHAN_CONTINUUM)
        while (G < 8 && G * 2 <= p.n_chan) G *= 2;
    const long long per_seg = (long long)p.n_baseline * ceil_div(p.n_chan, G);
    long long n_seg = ceil_div((long long)sm_count() * 2048 * 2, per_seg);   // ~2 waves of full occupancy
    if (n_seg < 1) n_seg = 1;
    p.seg_len = (int)ceil_div(p.n_time, n_seg);
    if (p.seg_len < 16) p.seg_len = 16;
    p.n_seg = (int)ceil_div(p.n_time, p.seg_len);
    const long long blocks = ceil_div(per_seg * p.n_seg, 256);
    CNGI_REQUIRE(blocks < (1LL << 31), "imaging_weight_grid: too many work items");
    cudaStream_t st = (cudaStream_t)stream;
    double *scale = nullptr;
    {
        int rc = make_uv_scale_table(p.freq, p.n_chan, p.dl, p.dm, p.n_u, p.n_v, st, &scale);
        if (rc != CNGI_OK) return rc;
        p.scale = scale;
    }
    // compile-time pol count when the pol_map is the identity (the reference's wrappers always pass arange)
    const int np = (!a->pol_map && (p.n_pol == 1 || p.n_pol == 2)) ? p.n_pol : 0;
    // fast kernel: whole channel groups, and every vector load it issues is aligned (two channels x 2 pols per 128-bit load in
    // fp32, one pol pair per load otherwise)
    const size_t elem = a->precision == CNGI_F32 ? 4 : 8;
    const size_t align = (np == 2) ? ((elem == 4 && G >= 2) ? 16 : 2 * elem) : elem;
    const bool fast = np && (p.n_chan % G == 0) && (reinterpret_cast<uintptr_t>(p.weight) % align == 0) &&
                      (((size_t)p.n_chan * np * elem) % align == 0) && (((size_t)G * np * elem) % align == 0) &&
                      p.n_chan <= 2048 && !env_flag("CNGI_IW_GRID_OLD");
    const size_t fast_smem = (size_t)2 * p.n_chan * sizeof(double);   // the transposed uv_scale table
#define CNGI_IW_LAUNCH(TT, GG)                                                      \
    do {                                                                            \
        if (fast && np == 2) iw_grid_fast_kernel<TT, GG, 2><<<(unsigned)blocks, 256, fast_smem, st>>>(p); \
        else if (fast) iw_grid_fast_kernel<TT, GG, 1><<<(unsigned)blocks, 256, fast_smem, st>>>(p); \
        else if (np == 2) iw_grid_kernel<TT, GG, 2><<<(unsigned)blocks, 256, 0, st>>>(p); \
        else if (np == 1) iw_grid_kernel<TT, GG, 1><<<(unsigned)blocks, 256, 0, st>>>(p); \
        else iw_grid_kernel<TT, GG, 0><<<(unsigned)blocks, 256, 0, st>>>(p);          \
    } while (0)
    if (a->precision == CNGI_F32) {
        switch (G) {
            case 1: CNGI_IW_LAUNCH(float, 1); break;
            case 2: CNGI_IW_LAUNCH(float, 2); break;
            case 4: CNGI_IW_LAUNCH(float, 4); break;
            default: CNGI_IW_LAUNCH(float, 8); break;
        }
    } else {
        switch (G) {
            case 1: CNGI_IW_LAUNCH(double, 1); break;
            case 2: CNGI_IW_LAUNCH(double, 2); break;
            case 4: CNGI_IW_LAUNCH(double, 4); break;
            default: CNGI_IW_LAUNCH(double, 8); break;
        }
    }
#undef CNGI_IW_LAUNCH
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(scale, st);
    CNGI_CUDA_TRY(e);
    return CNGI_OK;
}

extern "C" int cngi_b200_briggs_factors(const double *density, const double *sum_weight, double *briggs_factors,
                                        int64_t n_planes, int64_t n_cells, double robust, int32_t weighting, void *stream)
{
    using namespace cngi;
    CNGI_REQUIRE(density && sum_weight && briggs_factors, "briggs_factors: null pointer");
    CNGI_REQUIRE(n_planes > 0 && n_planes < 65536 && n_cells > 0, "briggs_factors: bad sizes");
    CNGI_REQUIRE(weighting == 0 || weighting == 1, "briggs_factors: weighting must be 0 (briggs) or 1 (uniform)");
    cudaStream_t st = (cudaStream_t)stream;
    if (weighting == 0) {
        CNGI_CUDA_TRY(cudaMemsetAsync(briggs_factors, 0, n_planes * sizeof(double), st));
        long long bx = ceil_div(n_cells, 256 * 8);
        const long long cap = ceil_div((long long)sm_count() * 8, n_planes);
        if (bx > cap) bx = cap;
        if (bx < 1) bx = 1;
        iw_sumsq_kernel<<<dim3((unsigned)bx, (unsigned)n_planes), 256, 0, st>>>(density, briggs_factors, n_cells);
        CNGI_CUDA_TRY(cudaGetLastError());
    }
    const double k = 5.0 * pow(10.0, -robust);
    iw_briggs_finalize_kernel<<<(unsigned)ceil_div(n_planes, 128), 128, 0, st>>>(briggs_factors, sum_weight, n_planes, k * k,
                                                                                weighting);
    CNGI_CUDA_TRY(cudaGetLastError());
    return CNGI_OK;
}

extern "C" int cngi_b200_imaging_weight_degrid(const cngi_iw_degrid_args *a, void *stream)
{
    using namespace cngi;
    CNGI_REQUIRE(a != nullptr, "imaging_weight_degrid: null args");
    CNGI_REQUIRE(a->natural_weight && a->uvw && a->freq_chan && a->density && a->briggs_factors && a->imaging_weight,
                 "imaging_weight_degrid: null array pointer");
    CNGI_REQUIRE(a->precision == CNGI_F32 || a->precision == CNGI_F64, "imaging_weight_degrid: bad precision");
    CNGI_REQUIRE(a->chan_mode != CNGI_CHAN_GENERAL || a->chan_map, "imaging_weight_degrid: chan_map is null");
    const long long total = (long long)a->n_time * a->n_baseline * a->n_chan;
    if (total == 0 || a->n_pol == 0) return CNGI_OK;
    IwParams p{};
    p.n_time = (int)a->n_time, p.n_baseline = (int)a->n_baseline, p.n_chan = (int)a->n_chan, p.n_pol = (int)a->n_pol;
    p.n_ic = (int)a->n_imag_chan, p.n_ip = (int)a->n_imag_pol, p.n_u = (int)a->n_u, p.n_v = (int)a->n_v;
    p.weight = a->natural_weight, p.uvw = a->uvw, p.freq = a->freq_chan, p.chan_map = a->chan_map, p.pol_map = a->pol_map;
    p.density = const_cast<double *>(a->density), p.dl = a->delta_lm[0], p.dm = a->delta_lm[1], p.chan_mode = a->chan_mode;
    p.bf = a->briggs_factors, p.out = a->imaging_weight;
    p.ds_u = a->density_stride[0], p.ds_v = a->density_stride[1], p.ds_c = a->density_stride[2], p.ds_p = a->density_stride[3];
    p.pol_shared = a->pol_shared != 0 && p.n_pol == 2;
    int cx = 1;
    while (cx < 128 && cx < p.n_chan) cx <<= 1;   // channels per block (power of two), 256 / cx rows per block
    const dim3 block(cx, 256 / cx);
    const long long rows = (long long)p.n_time * p.n_baseline;
    const long long gx = ceil_div(rows, block.y), gy = ceil_div(p.n_chan, cx);
    CNGI_REQUIRE(gx < (1LL << 31) && gy < 65536, "imaging_weight_degrid: too many samples");
    const dim3 grid((unsigned)gx, (unsigned)gy);
    const int np = (!a->pol_map && (p.n_pol == 1 || p.n_pol == 2)) ? p.n_pol : 0;
    cudaStream_t st = (cudaStream_t)stream;
    double *scale = nullptr;
    int rc = make_uv_scale_table(p.freq, p.n_chan, p.dl, p.dm, p.n_u, p.n_v, st, &scale);
    if (rc != CNGI_OK) return rc;
    p.scale = scale;
#ifndef CNGI_IW_DEGRID_ROWS
#define CNGI_IW_DEGRID_ROWS 4
#endif
    constexpr int kRows = CNGI_IW_DEGRID_ROWS;   // samples per thread in the fast path
    const dim3 grid_mlp((unsigned)ceil_div(rows, (long long)block.y * kRows), (unsigned)gy);
    // 32-bit cell indices: every (u, v, pol) offset inside one imaging channel's planes must fit an int
    const long long span = (long long)(p.n_u - 1) * p.ds_u + (long long)(p.n_v - 1) * p.ds_v + (long long)(p.n_pol - 1) * p.ds_p;
    const bool fast = np && p.ds_u >= 0 && p.ds_v >= 0 && p.ds_p >= 0 && span < (1LL << 31) && p.n_u > 0 && p.n_v > 0 &&
                      !env_flag("CNGI_IW_DEGRID_MLP");
#define CNGI_DG_LAUNCH(TT)                                                                      \
    do {                                                                                        \
        if (fast && np == 2 && p.pol_shared) iw_degrid_fast_kernel<TT, 2, kRows, true><<<grid_mlp, block, 0, st>>>(p);  \
        else if (fast && np == 2) iw_degrid_fast_kernel<TT, 2, kRows, false><<<grid_mlp, block, 0, st>>>(p);            \
        else if (fast) iw_degrid_fast_kernel<TT, 1, kRows, false><<<grid_mlp, block, 0, st>>>(p);                       \
        else if (np == 2) iw_degrid_mlp_kernel<TT, 2, kRows><<<grid_mlp, block, 0, st>>>(p);    \
        else if (np == 1) iw_degrid_mlp_kernel<TT, 1, kRows><<<grid_mlp, block, 0, st>>>(p);    \
        else iw_degrid_kernel<TT, 0><<<grid, block, 0, st>>>(p);                                \
    } while (0)
    if (a->precision == CNGI_F32)
        CNGI_DG_LAUNCH(float);
    else
        CNGI_DG_LAUNCH(double);
#undef CNGI_DG_LAUNCH
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(scale, st);
    CNGI_CUDA_TRY(e);
    return CNGI_OK;
}
