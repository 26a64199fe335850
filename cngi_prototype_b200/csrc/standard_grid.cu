// standard_grid.cu -- A1/A2 of SURVEY.md section 8: the prolate-spheroidal convolutional gridder.
// Replaces _standard_grid_jit (/root/reference/ngcasa/imaging/_imaging_utils/_standard_grid.py:242-371).
//
// Two kernels:
//
//  * std_grid_naive_kernel  -- one thread per (time, baseline, chan) sample, S*S global reductions per
//    polarisation.  Order-independent, any support; it is the correctness anchor and the fallback for
//    supports the track kernel is not instantiated for.
//
//  * std_grid_track_kernel  -- the product kernel.  A work item walks ONE baseline through time (and
//    through `G` neighbouring channels when they share an image plane) and keeps the S x S stamp it is
//    currently under in REGISTERS: S lanes per item, lane r owns the grid column u == r (mod S) and holds
//    S accumulators (v == j (mod S)) per polarisation.  Because a baseline moves a fraction of a cell per
//    integration, a sample usually lands on the same S x S cells as the previous one, so accumulators
//    are flushed to the grid (native REDG.ADD.F32x2 / F64) only when their cell leaves the stamp --
//    typically a few reductions per sample instead of S*S.  There is no shared-memory fp atomic anywhere:
//    on sm_100a those compile to ATOMS.CAST.SPIN loops (checked with cuobjdump), REDG is native.
//
//    Each warp runs a two-phase loop over rounds of 32 samples:
//      phase 1 (32 lanes, one sample each): coalesced vectorised loads of vis/weight/flag (+ uvw),
//        fp64 bit-exact cell/offset/mask math, tap lookup from the shared-memory CF table, and staging of
//        {cell ids, weighted data, the S u-taps and S v-taps in residue order} into the warp's slice of
//        shared memory;
//      phase 2 (S lanes per item, IPW items per warp): each item consumes its staged samples in order,
//        broadcast-reading the records, flushing on cell change, then S*PP FMAs per lane.
//    Only __syncwarp() separates the phases; warps never wait for each other.
#include "standard_grid.cuh"
#include <cstdlib>
#include <algorithm>

namespace cngi {

// ------------------------------------------------------------------------------------------------
//  naive kernel
// ------------------------------------------------------------------------------------------------
template <typename T, bool CPLX>
__global__ void __launch_bounds__(256) std_grid_naive_kernel(StdParams p)
{
    using CT = typename Cplx<T>::type;
    const long long total = (long long)p.n_time * p.n_baseline * p.n_chan;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = idx < total;
    const int c = in_range ? (int)(idx % p.n_chan) : 0;
    const long long tb = in_range ? idx / p.n_chan : 0;
    const int half = p.support / 2;
    CellPos cp;
    bool ok = in_range;
    if (ok) {
        ok = locate_centre(p.uvw[tb * 3], p.uvw[tb * 3 + 1], p.scale[c], p.scale[p.n_chan + c], p.n_u, p.n_v, cp);
    }
    if (ok) ok = stamp_inside(cp.uc, cp.vc, half, p.n_u, p.n_v);
    int uoff = 0, voff = 0, a_chan = 0;
    if (ok) {
        uoff = oversample_offset(cp.uc, cp.u_pos, p.oversampling);
        voff = oversample_offset(cp.vc, cp.v_pos, p.oversampling);
        a_chan = chan_of(p, c);
    }
    for (int ip = 0; ip < p.n_pol; ++ip) {   // uniform trip count: the warp-level reduction below needs it
        bool use = ok;
        double w = 0.0, wre = 0.0, wim = 0.0;
        int a_pol = 0;
        if (use) {
            const long long s = idx * p.n_pol + ip;
            w = (double)((const T *)p.weight)[s];
            if (p.do_psf) {
                wre = w;
            } else {
                const CT d = ((const CT *)p.vis)[s];
                weighted_vis((double)d.x, (double)d.y, w, wre, wim);
                if (p.flag && p.flag[s]) wre = nan("");
            }
            use = !masked(wre, wim);
        }
        double norm = 0.0;
        if (use) {
            a_pol = pol_of(p, ip);
            const long long plane = ((long long)a_chan * p.n_ip + a_pol) * p.n_u;
            for (int iv = -half; iv < p.support - half; ++iv) {
                const double cv = p.cgk[abs(p.oversampling * iv + voff)];
                for (int iu = -half; iu < p.support - half; ++iu) {
                    const double conv = p.cgk[abs(p.oversampling * iu + uoff)] * cv;
                    const long long cell = (plane + cp.uc + iu) * p.n_v + cp.vc + iv;
                    if (CPLX) {
                        CT val;
                        val.x = (T)(conv * wre);
                        val.y = (T)(conv * wim);
                        red_add((CT *)p.grid + cell, val);
                    } else {
                        red_add((T *)p.grid + cell, (T)(conv * wre));
                    }
                    norm += conv;
                }
            }
        }
        warp_grouped_add(p.sum_weight, a_chan * p.n_ip + a_pol, w * norm, use);
    }
}

// ------------------------------------------------------------------------------------------------
//  track kernel
// ------------------------------------------------------------------------------------------------
template <typename T, bool CPLX, int S, int PP> struct TrackCfg {
    // The register window is W x W cells, W = the power of two above the support (8 for S = 7): one spare column /
    // row of hysteresis, so a stamp that jitters by a cell between samples (channels swept back and forth across a
    // cell boundary) flushes nothing, and all residue arithmetic is a bit mask.
    static constexpr int W = (S < 4) ? 4 : (S < 8) ? 8 : (S < 16) ? 16 : 32;
    static constexpr int IPW = 32 / W;                                            // items per warp
    static constexpr int ITER = 32 / IPW;                                         // samples per item per round (== W)
    static constexpr int NV = CPLX ? PP : (PP + 1) / 2;                           // accumulator pairs per cell
    static constexpr int TPV = 16 / (int)sizeof(T);                               // T's per 16-byte vector
    static constexpr int WD = (NV * 2 + TPV - 1) / TPV * TPV;                     // padded weighted-data count
    static constexpr int OFF_IDX = 0;
    static constexpr int OFF_WD = 16;
    static constexpr int OFF_CV = OFF_WD + WD * (int)sizeof(T);
    static constexpr int OFF_CU = OFF_CV + W * (int)sizeof(T);
    static constexpr int RAW_BYTES = OFF_CU + W * (int)sizeof(T);
    // 16 * odd bytes per record: a quarter warp of 128-bit stores then covers all 32 banks.
    static constexpr int REC_BYTES = ((RAW_BYTES / 16) % 2 == 0) ? RAW_BYTES + 16 : RAW_BYTES;
    static constexpr int WARP_BYTES = 32 * REC_BYTES;
};

constexpr int kSameFlag = 1;   // record idx.w: this sample has the same (plane, uc, vc) as the item's previous one

#ifndef CNGI_TRACK_MINB128
#define CNGI_TRACK_MINB128 4   // min resident blocks per SM asked of ptxas (caps registers/thread); tuned on B200
#endif
#ifndef CNGI_TRACK_MINB256
#define CNGI_TRACK_MINB256 2
#endif

template <typename T, bool CPLX, int S, int PP, int BLK>
__global__ void __launch_bounds__(BLK, (sizeof(T) == 4 ? (BLK == 128 ? CNGI_TRACK_MINB128 : CNGI_TRACK_MINB256) : 1))
std_grid_track_kernel(StdParams p)
{
    using Cfg = TrackCfg<T, CPLX, S, PP>;
    using CT = typename Cplx<T>::type;
    using P2 = typename Pair<T>::type;
    constexpr int W = Cfg::W, IPW = Cfg::IPW, ITER = Cfg::ITER, NV = Cfg::NV;
    constexpr int HALF = S / 2;
    const unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(16) unsigned char smem[];
    T *table = reinterpret_cast<T *>(smem);
    const int table_bytes = (p.table_len * (int)sizeof(T) + 15) / 16 * 16;
    double *scale = reinterpret_cast<double *>(smem + table_bytes);   // uv_scale[0][c], uv_scale[1][c] of the window
    const int scale_bytes = 2 * p.c_n * (int)sizeof(double);
    // tapsum[off + os/2 + 1] = sum over the S taps of the stamp for oversampling offset `off`
    double *tapsum = reinterpret_cast<double *>(smem + table_bytes + scale_bytes);
    const int n_off = p.oversampling + 3;
    const int tapsum_bytes = (n_off * (int)sizeof(double) + 15) / 16 * 16;
    for (int i = threadIdx.x; i < p.table_len; i += blockDim.x) table[i] = (T)p.cgk[i];
    for (int i = threadIdx.x; i < p.c_n; i += blockDim.x) {
        const double f = p.freq[p.c_lo + i];
        scale[i] = uv_scale_of(f, p.dl, p.n_u);
        scale[p.c_n + i] = uv_scale_of(f, p.dm, p.n_v);
    }
    for (int i = threadIdx.x; i < n_off; i += blockDim.x) {
        const int off = i - p.oversampling / 2 - 1;
        double sum = 0.0;
        for (int q = 0; q < S; ++q) {
            const int k = abs(p.oversampling * (q - HALF) + off);
            if (k < p.table_len) sum += (double)(T)p.cgk[k];
        }
        tapsum[i] = sum;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long task = (long long)blockIdx.x * (BLK / 32) + warp;
    if (task >= p.n_tasks) return;   // no block-wide barrier after this point
    unsigned char *wbuf = smem + table_bytes + scale_bytes + tapsum_bytes + warp * Cfg::WARP_BYTES;

    // ---- task decode: (time segment, baseline, pol group, channel span), channel span fastest ------
    const int cspan = (int)(task % p.n_cspan);
    long long rest = task / p.n_cspan;
    const int pgrp = (int)(rest % p.n_pgrp);
    rest /= p.n_pgrp;
    const int b = (int)(rest % p.n_baseline);
    const int seg = (int)(rest / p.n_baseline);
    const int t_lo = seg * p.seg_len;
    const int t_hi = min(p.n_time, t_lo + p.seg_len);
    const int G = p.G;
    const int spr = ITER >> p.log2G;   // time steps per round
    const int c_base = p.c_lo + cspan * IPW * G;
    const int c_end = p.c_lo + p.c_n;
    const int p0 = pgrp * PP;
    const int npol = min(PP, p.n_pol - p0);
    const long long plane_cells = (long long)p.n_u * p.n_v;

    // ---- phase-1 role: lane <-> staged sample ------------------------------------------------------
    const int k1 = lane % IPW;
    const int q1 = lane / IPW;
    const int g1 = q1 & (G - 1);
    const int row1 = q1 >> p.log2G;
    // Channels of an item are walked boustrophedon (forward on even time steps, backward on odd ones) when they all
    // share one image plane, so that consecutive samples of an item are always uv neighbours.
    const bool zigzag = (p.chan_mode == CNGI_CHAN_CONTINUUM) && G > 1;
    const int c_fwd = c_base + k1 * G + g1;
    const int c_bwd = c_base + k1 * G + (G - 1 - g1);
    int c1 = c_fwd;                                  // channel of the sample held in the raw registers
    bool chan_ok = c_fwd < c_end;
    const bool any_chan_ok = zigzag ? (c_fwd < c_end || c_bwd < c_end) : chan_ok;
    const int a_chan1 = chan_ok ? chan_of(p, c_fwd) : 0;   // zigzag only runs in continuum mode: plane 0 either way
    double sw_acc[PP];
#pragma unroll
    for (int ip = 0; ip < PP; ++ip) sw_acc[ip] = 0.0;
    long long carry_key = -1;   // key of the item's last sample of the previous round (lanes < IPW use it)

    // ---- phase-2 role: lane <-> (item, u residue mod W) ----------------------------------------------
    const int k2 = lane / W;
    const int r2 = lane & (W - 1);
    int apol[PP];
#pragma unroll
    for (int ip = 0; ip < PP; ++ip) apol[ip] = (ip < npol) ? pol_of(p, p0 + ip) : 0;
    P2 acc[W][NV];
#pragma unroll
    for (int j = 0; j < W; ++j)
#pragma unroll
        for (int n = 0; n < NV; ++n) acc[j][n].x = acc[j][n].y = (T)0;
    // register window of the item: columns [lo_u, lo_u + W), rows [lo_v, lo_v + W) of plane cur_plane
    int cur_plane = -1, lo_u = 0, lo_v = 0;
    long long plane_off[PP];   // element offset of plane (cur_plane, apol[ip]) in the grid
#pragma unroll
    for (int ip = 0; ip < PP; ++ip) plane_off[ip] = 0;

    // flush accumulator j (grid cell (u, v) of the current plane) with native reductions.  Accumulators are cleared
    // only inside the non-zero test: unconditional writes in this divergent path make ptxas keep copies of them.
    auto flush_one = [&](int j, int u, int v) {
        const int cell = u * p.n_v + v;
        if (CPLX) {
#pragma unroll
            for (int ip = 0; ip < PP; ++ip) {
                if (ip < npol && (acc[j][ip].x != (T)0 || acc[j][ip].y != (T)0)) {
                    CT val;
                    val.x = acc[j][ip].x;
                    val.y = acc[j][ip].y;
                    red_add((CT *)p.grid + plane_off[ip] + cell, val);
                    acc[j][ip].x = acc[j][ip].y = (T)0;
                }
            }
        } else {
#pragma unroll
            for (int ip = 0; ip < PP; ++ip) {
                T &v1 = (ip & 1) ? acc[j][ip / 2].y : acc[j][ip / 2].x;
                if (ip < npol && v1 != (T)0) {
                    red_add((T *)p.grid + plane_off[ip] + cell, v1);
                    v1 = (T)0;
                }
            }
        }
    };
    auto my_column = [&]() { return lo_u + ((r2 - lo_u) & (W - 1)); };   // the u in the window with u == r2 (mod W)
    auto flush_column = [&]() {                                          // all W accumulators of this lane
        const int u = my_column();
#pragma unroll
        for (int j = 0; j < W; ++j) flush_one(j, u, lo_v + ((j - lo_v) & (W - 1)));
    };

    // ---- raw sample registers (software prefetch: loads of round n+1 fly during phase 2 of round n) --
    double raw_u = 0.0, raw_v = 0.0;
    CT raw_vis[PP];
    T raw_w[PP];
    unsigned raw_flag = 0;
    bool raw_ok = false;
    auto load_raw = [&](int t0) {
        const int t = t0 + row1;
        if (zigzag) {
            c1 = (t & 1) ? c_bwd : c_fwd;
            chan_ok = c1 < c_end;
        }
        raw_ok = chan_ok && (t < t_hi);
        raw_flag = 0;
        if (raw_ok) {
            const long long tb = (long long)t * p.n_baseline + b;
            raw_u = p.uvw[tb * 3];
            raw_v = p.uvw[tb * 3 + 1];
            const long long s = (tb * p.n_chan + c1) * p.n_pol + p0;
            if (PP == 2 && npol == 2 && (p.n_pol & 1) == 0) {   // 2 pols, aligned: one vector load each
                const T *wp = (const T *)p.weight + s;
                if (sizeof(T) == 4) {
                    const float2 w2 = *reinterpret_cast<const float2 *>(wp);
                    raw_w[0] = (T)w2.x;
                    raw_w[PP - 1] = (T)w2.y;
                } else {
                    const double2 w2 = *reinterpret_cast<const double2 *>(wp);
                    raw_w[0] = (T)w2.x;
                    raw_w[PP - 1] = (T)w2.y;
                }
                if (!p.do_psf) {
                    if (sizeof(T) == 4) {
                        const float4 d = *reinterpret_cast<const float4 *>((const CT *)p.vis + s);
                        raw_vis[0].x = (T)d.x;
                        raw_vis[0].y = (T)d.y;
                        raw_vis[PP - 1].x = (T)d.z;
                        raw_vis[PP - 1].y = (T)d.w;
                    } else {
                        raw_vis[0] = ((const CT *)p.vis)[s];
                        raw_vis[PP - 1] = ((const CT *)p.vis)[s + 1];
                    }
                    if (p.flag) {
                        const uchar2 f2 = *reinterpret_cast<const uchar2 *>(p.flag + s);
                        raw_flag = (f2.x ? 1u : 0u) | (f2.y ? 2u : 0u);
                    }
                }
            } else {
#pragma unroll
                for (int ip = 0; ip < PP; ++ip) {
                    if (ip < npol) {
                        raw_w[ip] = ((const T *)p.weight)[s + ip];
                        if (!p.do_psf) {
                            raw_vis[ip] = ((const CT *)p.vis)[s + ip];
                            if (p.flag && p.flag[s + ip]) raw_flag |= 1u << ip;
                        }
                    }
                }
            }
        }
    };

    // ---- phase 1: locate, mask, look up taps, stage ------------------------------------------------
    auto stage = [&]() {
        unsigned char *rec = wbuf + lane * Cfg::REC_BYTES;
        int4 idx = make_int4(-1, 0, 0, 0);
        long long key = -1;
        CellPos cp;
        bool ok = raw_ok;
        if (ok) ok = locate_centre(raw_u, raw_v, scale[c1 - p.c_lo], scale[p.c_n + c1 - p.c_lo], p.n_u, p.n_v, cp);
        if (ok) ok = stamp_inside(cp.uc, cp.vc, HALF, p.n_u, p.n_v);
        if (ok) {
            T wd[Cfg::WD];
#pragma unroll
            for (int i = 0; i < Cfg::WD; ++i) wd[i] = (T)0;
            double wsel[PP];
            bool any = false;
#pragma unroll
            for (int ip = 0; ip < PP; ++ip) {
                wsel[ip] = 0.0;
                if (ip < npol) {
                    const T w = raw_w[ip];
                    T wre = w, wim = (T)0;
                    bool use;
                    if (p.do_psf) {
                        use = !(isnan(w) || w == (T)0);
                    } else {
                        const T a = raw_vis[ip].x, bq = raw_vis[ip].y;
                        const bool flagged = (raw_flag >> ip) & 1u;
                        if (sizeof(T) == 4 && isfinite(a) && isfinite(bq) && isfinite(w)) {
                            // all finite: vis*w is NaN-free and is zero exactly when w == 0 or vis == 0, so the
                            // reference's mask (_standard_grid.py:340) can be evaluated without the fp64 products
                            use = !flagged && !(w == (T)0 || (a == (T)0 && bq == (T)0));
                            wre = a * w;
                            wim = bq * w;
                        } else {
                            double dre, dim;
                            weighted_vis((double)a, (double)bq, (double)w, dre, dim);
                            use = !flagged && !masked(dre, dim);
                            wre = (T)dre;
                            wim = (T)dim;
                        }
                    }
                    if (use) {
                        any = true;
                        wsel[ip] = (double)w;
                        if (CPLX) {
                            wd[2 * ip] = wre;
                            wd[2 * ip + 1] = wim;
                        } else {
                            wd[ip] = wre;   // pair n holds (pol 2n, pol 2n+1)
                        }
                    }
                }
            }
            if (any) {
                const int uoff = oversample_offset(cp.uc, cp.u_pos, p.oversampling);
                const int voff = oversample_offset(cp.vc, cp.v_pos, p.oversampling);
                T *rcu = reinterpret_cast<T *>(rec + Cfg::OFF_CU);
                T *rcv = reinterpret_cast<T *>(rec + Cfg::OFF_CV);
                // taps go to slot (cell mod W); the W - S slots outside the stamp get a zero tap
#pragma unroll
                for (int q = 0; q < W; ++q) {
                    T tu = (T)0, tv = (T)0;
                    if (q < S) {
                        tu = table[abs(p.oversampling * (q - HALF) + uoff)];
                        tv = table[abs(p.oversampling * (q - HALF) + voff)];
                    }
                    rcu[(cp.uc - HALF + q) & (W - 1)] = tu;
                    rcv[(cp.vc - HALF + q) & (W - 1)] = tv;
                }
                const int o0 = p.oversampling / 2 + 1;
                const double norm = tapsum[uoff + o0] * tapsum[voff + o0];   // == sum over the stamp of cu*cv
#pragma unroll
                for (int ip = 0; ip < PP; ++ip) sw_acc[ip] += wsel[ip] * norm;
                T *rwd = reinterpret_cast<T *>(rec + Cfg::OFF_WD);
                if (sizeof(T) == 4) {
#pragma unroll
                    for (int i = 0; i < Cfg::WD; i += 4)
                        *reinterpret_cast<float4 *>(rwd + i) = make_float4(wd[i], wd[i + 1], wd[i + 2], wd[i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < Cfg::WD; i += 2)
                        *reinterpret_cast<double2 *>(rwd + i) = make_double2(wd[i], wd[i + 1]);
                }
                idx = make_int4(cp.uc, cp.vc, a_chan1, 0);
                key = ((long long)a_chan1 * p.n_u + cp.uc) * p.n_v + cp.vc;
            }
        }
        // the item's previous sample is IPW slots back (or the last slot of the previous round)
        long long prev = __shfl_up_sync(FULL, key, IPW);
        if (lane < IPW) prev = carry_key;
        carry_key = __shfl_sync(FULL, key, 32 - IPW + k1);
        if (key >= 0 && key == prev) idx.w |= kSameFlag;
        *reinterpret_cast<int4 *>(rec + Cfg::OFF_IDX) = idx;
    };

    // ---- phase 2: consume ----------------------------------------------------------------------------
    auto consume = [&]() {
#pragma unroll 2
        for (int i = 0; i < ITER; ++i) {
            const unsigned char *rec = wbuf + (i * IPW + k2) * Cfg::REC_BYTES;
            const int4 idx = *reinterpret_cast<const int4 *>(rec + Cfg::OFF_IDX);
            // taps and data are fetched together with the cell ids (before the branches below), so an iteration
            // exposes one shared-memory latency instead of two
            const T cu = reinterpret_cast<const T *>(rec + Cfg::OFF_CU)[r2];
            T cv[W];
            P2 wd[Cfg::WD / 2];
            if (sizeof(T) == 4) {
#pragma unroll
                for (int q = 0; q < W; q += 4) {
                    const float4 x = *reinterpret_cast<const float4 *>(rec + Cfg::OFF_CV + q * 4);
                    cv[q] = x.x, cv[q + 1] = x.y, cv[q + 2] = x.z, cv[q + 3] = x.w;
                }
#pragma unroll
                for (int q = 0; q < Cfg::WD; q += 4) {
                    const float4 x = *reinterpret_cast<const float4 *>(rec + Cfg::OFF_WD + q * 4);
                    wd[q / 2].x = x.x, wd[q / 2].y = x.y, wd[q / 2 + 1].x = x.z, wd[q / 2 + 1].y = x.w;
                }
            } else {
#pragma unroll
                for (int q = 0; q < W; q += 2) {
                    const double2 x = *reinterpret_cast<const double2 *>(rec + Cfg::OFF_CV + q * 8);
                    cv[q] = x.x, cv[q + 1] = x.y;
                }
#pragma unroll
                for (int q = 0; q < Cfg::WD; q += 2) {
                    const double2 x = *reinterpret_cast<const double2 *>(rec + Cfg::OFF_WD + q * 8);
                    wd[q / 2].x = x.x, wd[q / 2].y = x.y;
                }
            }
            if (idx.x < 0) continue;
            if (!(idx.w & kSameFlag)) {   // the stamp moved (or first sample): does it still fit the register window?
                const int need_u = idx.x - HALF, need_v = idx.y - HALF;   // lowest column / row the stamp touches
                if (idx.z != cur_plane) {
                    if (cur_plane >= 0) flush_column();
#pragma unroll
                    for (int ip = 0; ip < PP; ++ip) plane_off[ip] = ((long long)idx.z * p.n_ip + apol[ip]) * plane_cells;
                    cur_plane = idx.z, lo_u = need_u, lo_v = need_v;
                } else {
                    // slide the window by the least amount that makes the stamp fit (hysteresis of W - S cells)
                    int new_u = lo_u, new_v = lo_v;
                    if (need_u < lo_u) new_u = need_u;
                    else if (need_u + S > lo_u + W) new_u = need_u + S - W;
                    if (need_v < lo_v) new_v = need_v;
                    else if (need_v + S > lo_v + W) new_v = need_v + S - W;
                    if (new_u != lo_u) {   // my column leaves iff it is outside the new column range
                        const int u = my_column();
                        if (u < new_u || u >= new_u + W) flush_column();
                        lo_u = new_u;
                    }
                    if (new_v != lo_v) {   // rows outside the new row range leave (cleared already if the column went)
                        const int u = my_column();
#pragma unroll
                        for (int j = 0; j < W; ++j) {
                            const int v = lo_v + ((j - lo_v) & (W - 1));
                            if (v < new_v || v >= new_v + W) flush_one(j, u, v);
                        }
                        lo_v = new_v;
                    }
                }
            }
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                const P2 t = pk_mul(wd[n], cu);
#pragma unroll
                for (int j = 0; j < W; ++j) pk_fma_acc(acc[j][n], t, cv[j]);
            }
        }
    };

    // ---- main loop over rounds ---------------------------------------------------------------------
    load_raw(t_lo);
    for (int t0 = t_lo; t0 < t_hi; t0 += spr) {
        stage();
        __syncwarp();
        if (t0 + spr < t_hi) load_raw(t0 + spr);
        consume();
        __syncwarp();
    }
    if (cur_plane >= 0) flush_column();

    // ---- sum_weight: lanes that share a channel reduce first, then one reduction per image plane -------
    const int span = IPW * G;   // lanes L and L + span handle the same channel
#pragma unroll
    for (int ip = 0; ip < PP; ++ip) {
        double v = sw_acc[ip];
        for (int o = span; o < 32; o <<= 1) v += __shfl_xor_sync(FULL, v, o);
        const bool lead = (lane < span) && any_chan_ok && (ip < npol);
        warp_grouped_add(p.sum_weight, a_chan1 * p.n_ip + apol[ip], v, lead);
    }
}

// ------------------------------------------------------------------------------------------------
//  host launchers
// ------------------------------------------------------------------------------------------------
static int validate(const cngi_std_grid_args *a)
{
    CNGI_REQUIRE(a != nullptr, "standard_grid: null args");
    CNGI_REQUIRE(a->n_time >= 0 && a->n_baseline >= 0 && a->n_chan >= 0 && a->n_pol >= 0,
                 "standard_grid: negative sample dimension");
    CNGI_REQUIRE(a->n_u > 0 && a->n_v > 0 && a->n_imag_chan > 0 && a->n_imag_pol > 0, "standard_grid: empty grid");
    CNGI_REQUIRE(a->n_u < (1 << 24) && a->n_v < (1 << 24), "standard_grid: grid side too large");
    CNGI_REQUIRE(a->n_u * a->n_v < (1LL << 31), "standard_grid: n_u*n_v overflows int32");
    CNGI_REQUIRE(a->n_time * a->n_baseline < (1LL << 31) && a->n_chan < (1 << 24) && a->n_pol <= 64,
                 "standard_grid: sample dimensions out of range");
    CNGI_REQUIRE(a->precision == CNGI_F32 || a->precision == CNGI_F64, "standard_grid: bad precision %d", a->precision);
    CNGI_REQUIRE(a->support >= 1 && a->oversampling >= 0, "standard_grid: bad support/oversampling");
    CNGI_REQUIRE(a->do_psf || a->complex_grid, "standard_grid: image mode needs a complex grid");
    CNGI_REQUIRE(a->do_psf || a->vis != nullptr, "standard_grid: vis is null in image mode");
    CNGI_REQUIRE(a->weight && a->uvw && a->freq_chan && a->cgk_1D && a->grid && a->sum_weight,
                 "standard_grid: null array pointer");
    CNGI_REQUIRE(a->chan_mode != CNGI_CHAN_GENERAL || a->chan_map != nullptr, "standard_grid: chan_map is null");
    CNGI_REQUIRE(a->chan_mode >= 0 && a->chan_mode <= 2, "standard_grid: bad chan_mode");
    if (a->chan_mode == CNGI_CHAN_CUBE)
        CNGI_REQUIRE(a->n_imag_chan >= a->n_chan, "standard_grid: cube mode needs n_imag_chan >= n_chan");
    if (!a->pol_map) CNGI_REQUIRE(a->n_imag_pol >= a->n_pol, "standard_grid: identity pol_map needs n_imag_pol >= n_pol");
    return CNGI_OK;
}

static StdParams make_params(const cngi_std_grid_args *a)
{
    StdParams p{};
    p.n_time = (int)a->n_time, p.n_baseline = (int)a->n_baseline, p.n_chan = (int)a->n_chan, p.n_pol = (int)a->n_pol;
    p.n_ic = (int)a->n_imag_chan, p.n_ip = (int)a->n_imag_pol, p.n_u = (int)a->n_u, p.n_v = (int)a->n_v;
    p.vis = a->vis, p.weight = a->weight, p.flag = a->do_psf ? nullptr : a->flag, p.uvw = a->uvw, p.freq = a->freq_chan;
    p.chan_map = a->chan_map, p.pol_map = a->pol_map, p.cgk = a->cgk_1D, p.grid = a->grid, p.sum_weight = a->sum_weight;
    p.dl = a->delta_lm[0], p.dm = a->delta_lm[1];
    p.support = a->support, p.oversampling = a->oversampling, p.do_psf = a->do_psf, p.chan_mode = a->chan_mode;
    p.table_len = a->oversampling * (a->support / 2 + 1);
    if (p.table_len < 1) p.table_len = 1;
    return p;
}

template <typename T, bool CPLX> static int launch_naive(StdParams p, cudaStream_t st)
{
    const long long total = (long long)p.n_time * p.n_baseline * p.n_chan;
    if (total == 0 || p.n_pol == 0) return CNGI_OK;
    const long long blocks = ceil_div(total, 256);
    CNGI_REQUIRE(blocks < (1LL << 31), "standard_grid: too many samples for one launch");
    double *scale = nullptr;
    int rc = make_uv_scale_table(p.freq, p.n_chan, p.dl, p.dm, p.n_u, p.n_v, st, &scale);
    if (rc != CNGI_OK) return rc;
    p.scale = scale;
    std_grid_naive_kernel<T, CPLX><<<(unsigned)blocks, 256, 0, st>>>(p);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(scale, st);
    CNGI_CUDA_TRY(e);
    return CNGI_OK;
}

static int track_block_override()
{
    // development knob: CNGI_TRACK_BLOCK=128|256 overrides the block size picked per precision
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("CNGI_TRACK_BLOCK");
        v = e ? atoi(e) : 0;
    }
    return v;
}

template <typename T, bool CPLX, int S, int PP, int BLK>
static int launch_track_blk(StdParams p, long long blocks, size_t smem, cudaStream_t st)
{
    auto kern = std_grid_track_kernel<T, CPLX, S, PP, BLK>;
    CNGI_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, BLK, smem, st>>>(p);
    CNGI_CUDA_TRY(cudaGetLastError());
    return CNGI_OK;
}

template <typename T, bool CPLX, int S, int PP>
static int launch_track(StdParams p, const cngi_std_grid_args *a, cudaStream_t st)
{
    using Cfg = TrackCfg<T, CPLX, S, PP>;
    if (p.n_time == 0 || p.n_baseline == 0 || p.n_chan == 0 || p.n_pol == 0) return CNGI_OK;
    constexpr int kMaxChanWindow = 2048;   // 32 KB of uv-scale table per block at most
    for (int c_lo = 0; c_lo < p.n_chan; c_lo += kMaxChanWindow) {
        p.c_lo = c_lo;
        p.c_n = std::min(kMaxChanWindow, p.n_chan - c_lo);
        // channels walked per item: only worth it when neighbouring channels share an image plane
        int G = a->chan_group;
        if (G <= 0) G = (p.chan_mode == CNGI_CHAN_CONTINUUM) ? Cfg::ITER : 1;
        if (G > Cfg::ITER) G = Cfg::ITER;
        while (G > 1 && (Cfg::IPW * G / 2) >= p.c_n) G >>= 1;   // do not span more channels than exist
        int log2G = 0;
        while ((1 << (log2G + 1)) <= G) ++log2G;
        G = 1 << log2G;
        p.G = G, p.log2G = log2G;
        const int spr = Cfg::ITER / G;
        p.n_cspan = (int)ceil_div(p.c_n, Cfg::IPW * G);
        p.n_pgrp = (int)ceil_div(p.n_pol, PP);
        const long long per_seg = (long long)p.n_baseline * p.n_cspan * p.n_pgrp;
        int seg_len = a->time_segment;
        if (seg_len <= 0) {
            // aim for ~16 resident-warp waves so the tail is small, but keep segments long enough that the
            // final flush (S*S cells per item) is amortised
            const long long target = (long long)sm_count() * 24 * 16;
            long long n_seg = ceil_div(target, per_seg);
            if (n_seg < 1) n_seg = 1;
            seg_len = (int)ceil_div(p.n_time, n_seg);
            const int min_len = 64 * spr / Cfg::ITER > 8 ? 64 * spr / Cfg::ITER : 8;
            if (seg_len < min_len) seg_len = min_len;
        }
        seg_len = (int)(ceil_div(seg_len, spr) * spr);
        p.seg_len = seg_len;
        p.n_seg = (int)ceil_div(p.n_time, seg_len);
        p.n_tasks = per_seg * p.n_seg;
        // fp64 accumulators need ~150 registers/thread: smaller blocks let three of them share an SM
        int blk = 128;
        if (track_block_override() == 128 || track_block_override() == 256) blk = track_block_override();
        const int wpb = blk / 32;
        const long long blocks = ceil_div(p.n_tasks, wpb);
        CNGI_REQUIRE(blocks < (1LL << 31), "standard_grid: too many work items for one launch");
        const size_t smem = (size_t)((p.table_len * (int)sizeof(T) + 15) / 16 * 16) + (size_t)2 * p.c_n * sizeof(double) +
                            (size_t)(((p.oversampling + 3) * (int)sizeof(double) + 15) / 16 * 16) +
                            (size_t)wpb * Cfg::WARP_BYTES;
        CNGI_REQUIRE(smem <= 227 * 1024, "standard_grid: CF table too large for shared memory (%zu bytes)", smem);
        int rc = blk == 128 ? launch_track_blk<T, CPLX, S, PP, 128>(p, blocks, smem, st)
                            : launch_track_blk<T, CPLX, S, PP, 256>(p, blocks, smem, st);
        if (rc != CNGI_OK) return rc;
    }
    return CNGI_OK;
}

template <typename T, bool CPLX, int S> static int launch_track_pp(StdParams p, const cngi_std_grid_args *a, cudaStream_t st)
{
    if (p.n_pol == 1) return launch_track<T, CPLX, S, 1>(p, a, st);
    return launch_track<T, CPLX, S, 2>(p, a, st);
}

template <typename T, bool CPLX> static int dispatch(StdParams p, const cngi_std_grid_args *a, cudaStream_t st)
{
#ifdef CNGI_TRACK_MINIMAL   // SASS experiments: instantiate one kernel only
    return launch_track<float, true, 7, 2>(p, a, st);
#else
    int algo = a->algorithm;
    const bool track_ok = (a->support == 3 || a->support == 5 || a->support == 7 || a->support == 9) &&
                          a->oversampling >= 1 && p.table_len <= 8192;
    const bool shift_ok = shift_kernel_supported(a, p.table_len);
    if (algo == CNGI_ALGO_AUTO) {
        // measured on B200 (tools/probe_std_grid.py): the window kernel wins every case tried
        // (fp32/fp64, continuum/cube, 1024^2 .. 8192^2); the shift and track kernels stay selectable
        if (window_kernel_supported(a, p.table_len)) algo = CNGI_ALGO_WINDOW;
        else algo = track_ok ? CNGI_ALGO_TRACK : CNGI_ALGO_NAIVE;
    }
    if (algo == CNGI_ALGO_WINDOW) {
        if (!window_kernel_supported(a, p.table_len)) {
            set_error("standard_grid: window kernel needs support in {3,5,7} and tap tables that fit shared memory");
            return CNGI_ERR_UNSUPPORTED;
        }
        return launch_window(p, a, st);
    }
    if (algo == CNGI_ALGO_SHIFT) {
        if (!shift_ok) {
            set_error("standard_grid: shift kernel needs support in {3,5,7,9} and tap tables that fit shared memory");
            return CNGI_ERR_UNSUPPORTED;
        }
        return launch_shift(p, a, st);
    }
    if (algo == CNGI_ALGO_TRACK) {
        if (!track_ok) {
            set_error("standard_grid: track kernel supports support in {3,5,7,9} (got %d)", a->support);
            return CNGI_ERR_UNSUPPORTED;
        }
        switch (a->support) {
            case 3: return launch_track_pp<T, CPLX, 3>(p, a, st);
            case 5: return launch_track_pp<T, CPLX, 5>(p, a, st);
            case 7: return launch_track_pp<T, CPLX, 7>(p, a, st);
            default: return launch_track_pp<T, CPLX, 9>(p, a, st);
        }
    }
    return launch_naive<T, CPLX>(p, st);
#endif
}

}  // namespace cngi

extern "C" int cngi_b200_standard_grid(const cngi_std_grid_args *a, void *stream)
{
    using namespace cngi;
    int rc = validate(a);
    if (rc != CNGI_OK) return rc;
    StdParams p = make_params(a);
    cudaStream_t st = (cudaStream_t)stream;
    if (a->precision == CNGI_F32)
        return a->complex_grid ? dispatch<float, true>(p, a, st) : dispatch<float, false>(p, a, st);
    return a->complex_grid ? dispatch<double, true>(p, a, st) : dispatch<double, false>(p, a, st);
}

// N1: the per-channel pipeline of synthesis_imaging_cube.py:195-211 grids the psf and the image of
// the same samples back to back; this entry point does both in one pass of the window kernel.
extern "C" int cngi_b200_standard_grid_image_psf(const cngi_std_grid_args *a, void *psf_grid, double *psf_sum_weight,
                                                 void *stream)
{
    using namespace cngi;
    int rc = validate(a);
    if (rc != CNGI_OK) return rc;
    CNGI_REQUIRE(psf_grid != nullptr && psf_sum_weight != nullptr, "standard_grid_image_psf: null psf outputs");
    CNGI_REQUIRE(!a->do_psf && a->complex_grid, "standard_grid_image_psf: args describe the image pass (do_psf 0, complex grid)");
    StdParams p = make_params(a);
    if (a->support != 7 || !window_kernel_supported(a, p.table_len)) {
        set_error("standard_grid_image_psf: the fused pass needs support 7 and tap tables that fit shared memory; "
                  "call cngi_b200_standard_grid twice instead");
        return CNGI_ERR_UNSUPPORTED;
    }
    p.psf_grid = psf_grid, p.psf_sum_weight = psf_sum_weight;
    return launch_window_dual(p, a, (cudaStream_t)stream);
}

// A4 folded into A1 (see include/cngi_b200.h): natural weights in, imaging weights formed in the gridder's phase 1.
extern "C" int cngi_b200_standard_grid_weighted(const cngi_std_grid_args *a, const cngi_iw_fused_args *w, void *stream)
{
    using namespace cngi;
    int rc = validate(a);
    if (rc != CNGI_OK) return rc;
    CNGI_REQUIRE(w != nullptr && w->density && w->briggs_factors, "standard_grid_weighted: null density / briggs_factors");
    CNGI_REQUIRE(!a->do_psf && a->complex_grid, "standard_grid_weighted: args describe the image pass (do_psf 0, complex grid)");
    CNGI_REQUIRE(w->n_u > 0 && w->n_v > 0 && w->n_u < (1 << 24) && w->n_v < (1 << 24), "standard_grid_weighted: bad density grid size");
    StdParams p = make_params(a);
    if (a->support != 7 || !window_kernel_supported(a, p.table_len)) {
        set_error("standard_grid_weighted: the fused pass needs support 7 and tap tables that fit shared memory; "
                  "call cngi_b200_imaging_weight_degrid and cngi_b200_standard_grid instead");
        return CNGI_ERR_UNSUPPORTED;
    }
    p.iw_density = w->density, p.iw_bf = w->briggs_factors, p.iw_out = w->imaging_weight;
    p.iw_ds_u = w->density_stride[0], p.iw_ds_v = w->density_stride[1];
    p.iw_ds_c = w->density_stride[2], p.iw_ds_p = w->density_stride[3];
    p.iw_n_u = (int)w->n_u, p.iw_n_v = (int)w->n_v, p.iw_dl = w->delta_lm[0], p.iw_dm = w->delta_lm[1];
    p.iw_pol_shared = w->pol_shared != 0;
    p.iw_own_scale = !(p.iw_n_u == p.n_u && p.iw_n_v == p.n_v && p.iw_dl == p.dl && p.iw_dm == p.dm);
    return launch_window_iw(p, a, (cudaStream_t)stream);
}
