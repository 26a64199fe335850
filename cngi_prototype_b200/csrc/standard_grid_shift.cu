// standard_grid_shift.cu -- A1 of SURVEY.md section 8: the product kernel of the prolate-spheroidal gridder.
// Replaces _standard_grid_jit (/root/reference/ngcasa/imaging/_imaging_utils/_standard_grid.py:242-371).
//
// Same decomposition as the "track" kernel in standard_grid.cu (a work item walks ONE baseline through a time
// segment, through G neighbouring channels when they share an image plane, and keeps the stamp it is under in
// registers; cells are reduced into the grid with native REDG only when they leave the register window), with the
// bookkeeping around the FMAs cut to the bone -- the track kernel issued 108 instructions per (4-sample) warp
// iteration of which 18 were packed FMAs (ncu, profiles/r01_std_grid_track_f32_continuum.txt):
//
//   * the register window is W columns x R = S + 1 rows.  Columns rotate (lane r of an item owns the column
//     u == r mod W, nothing moves when the window slides in u, the lane whose column left reduces its R cells).
//     Rows do NOT rotate: accumulator j is always row lo_v + j, and a slide in v reduces row 0 (or R - 1) and
//     SHIFTS the accumulators one register down (up).  Every accumulator index is therefore a compile-time
//     constant -- no unrolled "which of my W rows left" search, no per-cell address arithmetic.
//   * taps are never staged per sample.  Shared memory holds the tap rows pre-expanded per oversampling offset:
//     tapu[off][W] (S taps, zero padded) and tapv[dv][off][R] for the two places (dv = 0, 1) the stamp can sit in
//     the R-row window, so a lane fetches its u tap with one LDS.32 and its R v-taps with R/4 LDS.128 straight
//     from the table; phase 1 only stages {uc, vc, row byte offsets} + the weighted data (32 B per sample).
//   * the tap loads are issued before the "does the stamp still fit" branch (the slow path reloads them), and the
//     next record is fetched before the FMAs of the current one, so an iteration exposes one shared-memory latency.
//   * the boustrophedon channel walk permutes the staging SLOT instead of the lane's channel, so a lane streams
//     one channel for the whole segment: constant uv_scale registers, pointer-bump addressing.
#include "standard_grid.cuh"
#include <algorithm>

namespace cngi {

template <typename T, bool CPLX, int S, int PP> struct ShiftCfg {
    static constexpr int W = (S < 4) ? 4 : (S < 8) ? 8 : 16;      // lanes per item == columns of the register window
    static constexpr int R = S + 1;                              // rows of the register window (one spare row)
    static constexpr int IPW = 32 / W;                           // items per warp
    static constexpr int ITER = W;                               // samples per item per round
    static constexpr int NV = CPLX ? PP : (PP + 1) / 2;          // accumulator pairs per cell
    static constexpr int TPV = 16 / (int)sizeof(T);              // T's per 16-byte vector
    static constexpr int RP = (R + TPV - 1) / TPV * TPV;         // padded v-tap row
    static constexpr int WD = (NV * 2 + TPV - 1) / TPV * TPV;    // padded weighted-data count per record
    static constexpr int IDX_BYTES = 32 * 16;                    // 32 x int4 {uc, vc, u row byte offset, v row byte offset}
    static constexpr int WD_BYTES = 32 * WD * (int)sizeof(T);
    static constexpr int WARP_BYTES = IDX_BYTES + WD_BYTES;
};

struct ShiftSmem {
    int tapu, tapv, tapsum, scale, wbuf, total;
};

template <typename Cfg, typename T> __host__ __device__ inline ShiftSmem shift_smem_layout(int oversampling, int c_n, int warps)
{
    ShiftSmem L;
    const int n_off = oversampling + 3;
    auto up16 = [](int x) { return (x + 15) / 16 * 16; };
    L.tapu = 0;
    L.tapv = L.tapu + up16(n_off * Cfg::W * (int)sizeof(T));
    L.tapsum = L.tapv + up16(2 * n_off * Cfg::RP * (int)sizeof(T));
    L.scale = L.tapsum + up16(n_off * (int)sizeof(double));
    L.wbuf = L.scale + up16(2 * c_n * (int)sizeof(double));
    L.total = L.wbuf + warps * Cfg::WARP_BYTES;
    return L;
}

#ifndef CNGI_SHIFT_MINB_F32
#define CNGI_SHIFT_MINB_F32 4
#endif
#ifndef CNGI_SHIFT_MINB_F64
#define CNGI_SHIFT_MINB_F64 3
#endif
#ifndef CNGI_SHIFT_UNROLL
#define CNGI_SHIFT_UNROLL 2
#endif

template <typename T, bool CPLX, int S, int PP, int BLK>
__global__ void __launch_bounds__(BLK, (sizeof(T) == 4 ? CNGI_SHIFT_MINB_F32 : CNGI_SHIFT_MINB_F64))
std_grid_shift_kernel(StdParams p)
{
    using Cfg = ShiftCfg<T, CPLX, S, PP>;
    using CT = typename Cplx<T>::type;
    using P2 = typename Pair<T>::type;
    constexpr int W = Cfg::W, R = Cfg::R, IPW = Cfg::IPW, ITER = Cfg::ITER, NV = Cfg::NV, RP = Cfg::RP, WD = Cfg::WD;
    constexpr int HALF = S / 2;
    constexpr int SPARE_U = W - S;
    constexpr int kNoWindow = -(1 << 30);
    constexpr int kUnroll = CNGI_SHIFT_UNROLL;
    const unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(16) unsigned char smem[];
    const ShiftSmem L = shift_smem_layout<Cfg, T>(p.oversampling, p.c_n, BLK / 32);
    T *tapu = reinterpret_cast<T *>(smem + L.tapu);
    T *tapv = reinterpret_cast<T *>(smem + L.tapv);
    double *tapsum = reinterpret_cast<double *>(smem + L.tapsum);
    double *scale = reinterpret_cast<double *>(smem + L.scale);
    const int n_off = p.oversampling + 3;
    const int o0 = p.oversampling / 2 + 1;   // table row of oversampling offset 0
    auto tap_of = [&](int q, int off) -> T {  // tap q (0..S-1) of the stamp for oversampling offset `off`
        if (q < 0 || q >= S) return (T)0;
        const int k = abs(p.oversampling * (q - HALF) + off);
        return k < p.table_len ? (T)p.cgk[k] : (T)0;
    };
    for (int i = threadIdx.x; i < n_off * W; i += BLK) tapu[i] = tap_of(i % W, i / W - o0);
    for (int i = threadIdx.x; i < 2 * n_off * RP; i += BLK) {
        const int dv = i / (n_off * RP), rem = i % (n_off * RP);
        tapv[i] = tap_of(rem % RP - dv, rem / RP - o0);
    }
    for (int i = threadIdx.x; i < n_off; i += BLK) {
        double sum = 0.0;
        for (int q = 0; q < S; ++q) sum += (double)tap_of(q, i - o0);
        tapsum[i] = sum;
    }
    for (int i = threadIdx.x; i < p.c_n; i += BLK) {
        const double f = p.freq[p.c_lo + i];
        scale[i] = uv_scale_of(f, p.dl, p.n_u);
        scale[p.c_n + i] = uv_scale_of(f, p.dm, p.n_v);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long task = (long long)blockIdx.x * (BLK / 32) + warp;
    if (task >= p.n_tasks) return;   // no block-wide barrier after this point
    unsigned char *wbuf = smem + L.wbuf + warp * Cfg::WARP_BYTES;
    int4 *idx_arr = reinterpret_cast<int4 *>(wbuf);
    unsigned char *wd_arr = wbuf + Cfg::IDX_BYTES;

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

    // ---- phase-1 role: lane <-> one channel of one item, `row1`-th time step of the round ------------
    const int k1 = lane % IPW;
    const int q1 = lane / IPW;
    const int g1 = q1 & (G - 1);
    const int row1 = q1 >> p.log2G;
    const int c1 = c_base + k1 * G + g1;
    const bool chan_ok = c1 < c_end;
    const int a_chan1 = chan_ok ? chan_of(p, c1) : 0;
    const double su = chan_ok ? scale[c1 - p.c_lo] : 0.0;
    const double sv = chan_ok ? scale[p.c_n + c1 - p.c_lo] : 0.0;
    // The channels of an item are consumed boustrophedon (forward on even time steps, backward on odd ones) when
    // they share one image plane, so that consecutive samples of an item are always uv neighbours: the lane keeps
    // its channel and stages into the mirrored slot on odd time steps.
    const bool zigzag = (p.chan_mode == CNGI_CHAN_CONTINUUM) && G > 1;
    const int slot_fwd = lane;
    const int slot_bwd = zigzag ? ((row1 * G + (G - 1 - g1)) * IPW + k1) : lane;
    double sw_acc[PP];
#pragma unroll
    for (int ip = 0; ip < PP; ++ip) sw_acc[ip] = 0.0;

    // ---- phase-2 role: lane <-> (item, u residue mod W) ----------------------------------------------
    const int k2 = lane / W;
    const int r2 = lane & (W - 1);
    int apol[PP];
#pragma unroll
    for (int ip = 0; ip < PP; ++ip) apol[ip] = (ip < npol) ? pol_of(p, p0 + ip) : 0;
    const int c_item = c_base + k2 * G;   // every channel of an item maps to one image plane (G > 1 only in continuum)
    const int plane2 = (c_item < c_end) ? chan_of(p, c_item) : 0;
    T *gplane[PP];                        // plane (plane2, apol[ip]) of the grid, as T elements
#pragma unroll
    for (int ip = 0; ip < PP; ++ip)
        gplane[ip] = (T *)p.grid + ((long long)plane2 * p.n_ip + apol[ip]) * ((long long)p.n_u * p.n_v) * (CPLX ? 2 : 1);
    P2 acc[R][NV];
#pragma unroll
    for (int j = 0; j < R; ++j)
#pragma unroll
        for (int n = 0; n < NV; ++n) acc[j][n].x = acc[j][n].y = (T)0;
    int lo_u = kNoWindow, lo_v = 0;   // register window: columns [lo_u, lo_u + W), rows [lo_v, lo_v + R)
    const unsigned tapu_s = (unsigned)__cvta_generic_to_shared(tapu);
    const unsigned tapv_s = (unsigned)__cvta_generic_to_shared(tapv);
    const int vstride = n_off * RP * (int)sizeof(T);   // bytes between the dv = 0 and dv = 1 v-tap tables

    // reduce accumulator row j into grid cell `cell` of the item's planes (j is a constant after unrolling)
    auto flush_acc = [&](int j, int cell) {
        if (CPLX) {
#pragma unroll
            for (int ip = 0; ip < PP; ++ip) {
                if (ip < npol && (acc[j][ip].x != (T)0 || acc[j][ip].y != (T)0)) {
                    CT val;
                    val.x = acc[j][ip].x, val.y = acc[j][ip].y;
                    red_add((CT *)gplane[ip] + cell, val);
                }
            }
        } else {
#pragma unroll
            for (int ip = 0; ip < PP; ++ip) {
                const T v1 = (ip & 1) ? acc[j][ip / 2].y : acc[j][ip / 2].x;
                if (ip < npol && v1 != (T)0) red_add(gplane[ip] + cell, v1);
            }
        }
    };
    auto my_column = [&]() { return lo_u + ((r2 - lo_u) & (W - 1)); };   // the u in the window with u == r2 (mod W)
    auto clear_acc = [&]() {
#pragma unroll
        for (int j = 0; j < R; ++j)
#pragma unroll
            for (int n = 0; n < NV; ++n) acc[j][n].x = acc[j][n].y = (T)0;
    };
    // All R cells of this lane's column.  The spare row/columns of the window may lie outside the grid (their
    // accumulators only ever received zero taps): rows 0 and R-1 and the column are range checked.
    auto flush_column = [&]() {
        const int u = my_column();
        if ((unsigned)u < (unsigned)p.n_u) {
            const int cell0 = u * p.n_v + lo_v;
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const bool outside = (j == 0 && lo_v < 0) || (j == R - 1 && lo_v + R - 1 >= p.n_v);
                if (!outside) flush_acc(j, cell0 + j);
            }
        }
        clear_acc();
    };
    auto shift_up = [&]() {   // window moves one row up: row lo_v leaves through accumulator 0
        const int u = my_column();
        if ((unsigned)u < (unsigned)p.n_u && lo_v >= 0) flush_acc(0, u * p.n_v + lo_v);
#pragma unroll
        for (int j = 0; j < R - 1; ++j)
#pragma unroll
            for (int n = 0; n < NV; ++n) acc[j][n] = acc[j + 1][n];
#pragma unroll
        for (int n = 0; n < NV; ++n) acc[R - 1][n].x = acc[R - 1][n].y = (T)0;
        ++lo_v;
    };
    auto shift_down = [&]() {   // window moves one row down: row lo_v + R - 1 leaves through accumulator R - 1
        const int u = my_column();
        const int v = lo_v + R - 1;
        if ((unsigned)u < (unsigned)p.n_u && v < p.n_v) flush_acc(R - 1, u * p.n_v + v);
#pragma unroll
        for (int j = R - 1; j > 0; --j)
#pragma unroll
            for (int n = 0; n < NV; ++n) acc[j][n] = acc[j - 1][n];
#pragma unroll
        for (int n = 0; n < NV; ++n) acc[0][n].x = acc[0][n].y = (T)0;
        --lo_v;
    };
    // make the stamp whose lowest cell is (need_u, need_v) fit the window, sliding it by the least amount
    auto slide = [&](int need_u, int need_v) {
        if (lo_u == kNoWindow) {
            lo_u = need_u, lo_v = need_v;
            return;
        }
        const int du = need_u - lo_u;
        if ((unsigned)du > (unsigned)SPARE_U) {
            const int new_u = du < 0 ? need_u : need_u - SPARE_U;
            const int u = my_column();
            if (u < new_u || u >= new_u + W) flush_column();   // my column leaves
            lo_u = new_u;
        }
        const int dv = need_v - lo_v;
        if ((unsigned)dv > 1u) {
            const int new_v = dv < 0 ? need_v : need_v - 1;
            const int sh = new_v - lo_v;
            if (sh >= R || sh <= -R) {
                flush_column();
                lo_v = new_v;
            } else if (sh > 0) {
                for (int s = 0; s < sh; ++s) shift_up();
            } else {
                for (int s = 0; s < -sh; ++s) shift_down();
            }
        }
    };

    // ---- raw sample registers (software prefetch: loads of round n+1 fly during phase 2 of round n) --
    double raw_u = 0.0, raw_v = 0.0;
    CT raw_vis[PP];
    T raw_w[PP];
    unsigned raw_flag = 0;
    bool raw_ok = false;
    const long long s_step = (long long)spr * p.n_baseline * p.n_chan * p.n_pol;
    long long s_next = (((long long)(t_lo + row1) * p.n_baseline + b) * p.n_chan + c1) * p.n_pol + p0;
    const double *uvw_next = p.uvw + ((long long)(t_lo + row1) * p.n_baseline + b) * 3;
    const long long uvw_step = (long long)spr * p.n_baseline * 3;
    const bool vec2 = (PP == 2) && npol == 2 && (p.n_pol & 1) == 0;
    auto load_raw = [&](int t0) {
        raw_ok = chan_ok && (t0 + row1 < t_hi);
        raw_flag = 0;
        if (raw_ok) {
            raw_u = uvw_next[0];
            raw_v = uvw_next[1];
            const long long s = s_next;
            if (vec2) {   // 2 pols, aligned: one vector load each
                const T *wp = (const T *)p.weight + s;
                if (sizeof(T) == 4) {
                    const float2 w2 = *reinterpret_cast<const float2 *>(wp);
                    raw_w[0] = (T)w2.x, raw_w[PP - 1] = (T)w2.y;
                } else {
                    const double2 w2 = *reinterpret_cast<const double2 *>(wp);
                    raw_w[0] = (T)w2.x, raw_w[PP - 1] = (T)w2.y;
                }
                if (!p.do_psf) {
                    if (sizeof(T) == 4) {
                        const float4 d = *reinterpret_cast<const float4 *>((const CT *)p.vis + s);
                        raw_vis[0].x = (T)d.x, raw_vis[0].y = (T)d.y;
                        raw_vis[PP - 1].x = (T)d.z, raw_vis[PP - 1].y = (T)d.w;
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
        s_next += s_step;
        uvw_next += uvw_step;
    };

    // ---- phase 1: locate, mask, stage ----------------------------------------------------------------
    auto stage = [&](int t0) {
        int4 idx = make_int4(-1, 0, 0, 0);
        const int slot = ((t0 + row1) & 1) ? slot_bwd : slot_fwd;
        CellPos cp;
        bool ok = raw_ok;
        if (ok) ok = locate_centre(raw_u, raw_v, su, sv, p.n_u, p.n_v, cp);
        if (ok) ok = stamp_inside(cp.uc, cp.vc, HALF, p.n_u, p.n_v);
        if (ok) {
            T wd[WD];
#pragma unroll
            for (int i = 0; i < WD; ++i) wd[i] = (T)0;
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
                const int uo = oversample_offset(cp.uc, cp.u_pos, p.oversampling) + o0;
                const int vo = oversample_offset(cp.vc, cp.v_pos, p.oversampling) + o0;
                const double norm = tapsum[uo] * tapsum[vo];   // == sum over the stamp of cu*cv
#pragma unroll
                for (int ip = 0; ip < PP; ++ip) sw_acc[ip] += wsel[ip] * norm;
                T *rwd = reinterpret_cast<T *>(wd_arr + slot * (WD * (int)sizeof(T)));
                if (sizeof(T) == 4) {
#pragma unroll
                    for (int i = 0; i < WD; i += 4)
                        *reinterpret_cast<float4 *>(rwd + i) = make_float4(wd[i], wd[i + 1], wd[i + 2], wd[i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < WD; i += 2)
                        *reinterpret_cast<double2 *>(rwd + i) = make_double2(wd[i], wd[i + 1]);
                }
                idx = make_int4(cp.uc, cp.vc, uo * (W * (int)sizeof(T)), vo * (RP * (int)sizeof(T)));
            }
        }
        idx_arr[slot] = idx;
    };

    // ---- phase 2: consume ----------------------------------------------------------------------------
    auto lds_vec = [](unsigned addr, T *dst) {   // one 16-byte shared-memory load
        if constexpr (sizeof(T) == 4) {
            float4 x;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(addr));
            dst[0] = (T)x.x, dst[1] = (T)x.y, dst[2] = (T)x.z, dst[3] = (T)x.w;
        } else {
            double2 x;
            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x.x), "=d"(x.y) : "r"(addr));
            dst[0] = (T)x.x, dst[1] = (T)x.y;
        }
    };
    auto consume = [&]() {
        int4 nidx = idx_arr[k2];
        T nwd[WD];
        {
            const T *src = reinterpret_cast<const T *>(wd_arr + k2 * (WD * (int)sizeof(T)));
#pragma unroll
            for (int q = 0; q < WD; ++q) nwd[q] = src[q];
        }
#pragma unroll kUnroll
        for (int i = 0; i < ITER; ++i) {
            const int4 idx = nidx;
            T wd[WD];
#pragma unroll
            for (int q = 0; q < WD; ++q) wd[q] = nwd[q];
            // taps are fetched before the branches below (the slow path reloads the v taps)
            const int need_u = idx.x - HALF, need_v = idx.y - HALF;
            int dv = need_v - lo_v;
            const unsigned cu_addr = tapu_s + idx.z + (((r2 - need_u) & (W - 1)) * (int)sizeof(T));
            unsigned cv_addr = tapv_s + idx.w + ((dv & 1) ? vstride : 0);
            T cu;
            if constexpr (sizeof(T) == 4) {
                float x;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(cu_addr));
                cu = (T)x;
            } else {
                double x;
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(cu_addr));
                cu = (T)x;
            }
            T cv[RP];
#pragma unroll
            for (int q = 0; q < RP; q += Cfg::TPV) lds_vec(cv_addr + q * (int)sizeof(T), cv + q);
            if (i + 1 < ITER) {   // next record
                nidx = idx_arr[(i + 1) * IPW + k2];
                const T *src = reinterpret_cast<const T *>(wd_arr + ((i + 1) * IPW + k2) * (WD * (int)sizeof(T)));
#pragma unroll
                for (int q = 0; q < WD; ++q) nwd[q] = src[q];
            }
            if (idx.x < 0) continue;
            if ((unsigned)(need_u - lo_u) > (unsigned)SPARE_U || (unsigned)dv > 1u) {
                slide(need_u, need_v);
                dv = need_v - lo_v;
                cv_addr = tapv_s + idx.w + (dv ? vstride : 0);
#pragma unroll
                for (int q = 0; q < RP; q += Cfg::TPV) lds_vec(cv_addr + q * (int)sizeof(T), cv + q);
            }
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                P2 w2;
                w2.x = wd[2 * n], w2.y = wd[2 * n + 1];
                const P2 t = pk_mul(w2, cu);
#pragma unroll
                for (int j = 0; j < R; ++j) pk_fma_acc(acc[j][n], t, cv[j]);
            }
        }
    };

    // ---- main loop over rounds ---------------------------------------------------------------------
    load_raw(t_lo);
    for (int t0 = t_lo; t0 < t_hi; t0 += spr) {
        stage(t0);
        __syncwarp();
        if (t0 + spr < t_hi) load_raw(t0 + spr);
        consume();
        __syncwarp();
    }
    if (lo_u != kNoWindow) flush_column();

    // ---- sum_weight: lanes that share a channel reduce first, then one reduction per image plane -------
    const int span = IPW * G;   // lanes L and L + span handle the same channel
#pragma unroll
    for (int ip = 0; ip < PP; ++ip) {
        double v = sw_acc[ip];
        for (int o = span; o < 32; o <<= 1) v += __shfl_xor_sync(FULL, v, o);
        const bool lead = (lane < span) && chan_ok && (ip < npol);
        warp_grouped_add(p.sum_weight, a_chan1 * p.n_ip + apol[ip], v, lead);
    }
}

// ------------------------------------------------------------------------------------------------
//  host launcher
// ------------------------------------------------------------------------------------------------
template <typename T, bool CPLX, int S, int PP>
static int launch_shift_t(StdParams p, const cngi_std_grid_args *a, cudaStream_t st)
{
    using Cfg = ShiftCfg<T, CPLX, S, PP>;
    constexpr int BLK = 128;
    if (p.n_time == 0 || p.n_baseline == 0 || p.n_chan == 0 || p.n_pol == 0) return CNGI_OK;
    constexpr int kMaxChanWindow = 2048;   // 32 KB of uv-scale table per block at most
    auto kern = std_grid_shift_kernel<T, CPLX, S, PP, BLK>;
    for (int c_lo = 0; c_lo < p.n_chan; c_lo += kMaxChanWindow) {
        p.c_lo = c_lo;
        p.c_n = std::min(kMaxChanWindow, p.n_chan - c_lo);
        // channels walked per item: only worth it when neighbouring channels share an image plane
        int G = a->chan_group;
        if (G <= 0) G = (p.chan_mode == CNGI_CHAN_CONTINUUM) ? Cfg::ITER : 1;
        if (p.chan_mode != CNGI_CHAN_CONTINUUM) G = 1;   // an item owns ONE image plane
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
            // final flush (W*R cells per item) is amortised
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
        const long long blocks = ceil_div(p.n_tasks, BLK / 32);
        CNGI_REQUIRE(blocks < (1LL << 31), "standard_grid: too many work items for one launch");
        const size_t smem = (size_t)shift_smem_layout<Cfg, T>(p.oversampling, p.c_n, BLK / 32).total;
        CNGI_REQUIRE(smem <= 227 * 1024, "standard_grid: tap tables too large for shared memory (%zu bytes)", smem);
        CNGI_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)blocks, BLK, smem, st>>>(p);
        CNGI_CUDA_TRY(cudaGetLastError());
    }
    return CNGI_OK;
}

template <typename T, bool CPLX, int S> static int launch_shift_pp(StdParams p, const cngi_std_grid_args *a, cudaStream_t st)
{
    if (p.n_pol == 1) return launch_shift_t<T, CPLX, S, 1>(p, a, st);
    return launch_shift_t<T, CPLX, S, 2>(p, a, st);
}

template <typename T, bool CPLX> static int launch_shift_s(StdParams p, const cngi_std_grid_args *a, cudaStream_t st)
{
#ifdef CNGI_SHIFT_MINIMAL   // SASS experiments: instantiate one kernel only
    return launch_shift_t<float, true, 7, 2>(p, a, st);
#else
    switch (a->support) {
        case 3: return launch_shift_pp<T, CPLX, 3>(p, a, st);
        case 5: return launch_shift_pp<T, CPLX, 5>(p, a, st);
        case 7: return launch_shift_pp<T, CPLX, 7>(p, a, st);
        default: return launch_shift_pp<T, CPLX, 9>(p, a, st);
    }
#endif
}

bool shift_kernel_supported(const cngi_std_grid_args *a, int table_len)
{
    if (!(a->support == 3 || a->support == 5 || a->support == 7 || a->support == 9)) return false;
    if (a->oversampling < 1 || table_len > 8192) return false;
    // tap tables: (W + 2*RP) T's + one double per oversampling offset; keep them under ~64 KB so 3-4 blocks fit an SM
    const int n_off = a->oversampling + 3;
    const int tsz = a->precision == CNGI_F32 ? 4 : 8;
    const int w = a->support < 4 ? 4 : a->support < 8 ? 8 : 16;
    const int rp = (a->support + 1 + 3) / 4 * 4;
    return (long long)n_off * ((w + 2 * rp) * tsz + 8) <= 64 * 1024;
}

int launch_shift(StdParams p, const cngi_std_grid_args *a, cudaStream_t st)
{
    if (a->precision == CNGI_F32)
        return a->complex_grid ? launch_shift_s<float, true>(p, a, st) : launch_shift_s<float, false>(p, a, st);
    return a->complex_grid ? launch_shift_s<double, true>(p, a, st) : launch_shift_s<double, false>(p, a, st);
}

}  // namespace cngi
