// standard_grid_window.cuh -- A1 of SURVEY.md section 8: the product kernel of the prolate-spheroidal gridder.
// Replaces _standard_grid_jit (/root/reference/ngcasa/imaging/_imaging_utils/_standard_grid.py:242-371).
//
// Same decomposition as the "track" kernel in standard_grid.cu: a work item walks ONE baseline through a time
// segment (through G neighbouring channels when they share an image plane) and keeps the W x W cells it is under in
// REGISTERS -- W lanes per item, lane r owns the grid row v == r (mod W) (v is the contiguous axis of the grid),
// accumulator j is the column u == j (mod W) -- and cells are reduced into the grid with native REDG only when they
// leave that window.  What changed is
// everything around the FMAs; ncu on the track kernel (profiles/r01_std_grid_track_f32_continuum.txt) showed 108
// instructions per 4-sample warp iteration of which 18 were packed FMAs, and 300-instruction window slides:
//
//   * Taps are never staged per sample.  Shared memory holds every tap row the kernel can need, pre-rotated:
//     tap[rot][off][W] = the S taps for oversampling offset `off`, zero padded to W and rotated by `rot`, so a lane
//     fetches its v tap with one LDS.32 and its W u-taps (already in accumulator order) with W/4 LDS.128.  Phase 1
//     stages {lowest stamp cell, two row byte-offsets} + the weighted data: 32 B per sample, no tap arithmetic.
//   * Phase 2's fast path is ~25 instructions around the 18 packed FMAs of a sample: one packed compare decides
//     "the stamp still fits the window" (cell ids are staged as u<<16|v).  Consuming two samples per iteration was
//     measured too (NS = 2): no faster, the shared-memory pipe is the co-limiter.
//   * Window slides are cheap: a column leaving the window is reduced by all W lanes through a static binary
//     dispatch on (u mod W) -- 3 branches, 2 REDG covering W consecutive cells -- and a grid row leaving reduces one
//     lane's W accumulators; accumulators are only READ there and cleared afterwards with selects.
//   * Raw samples are prefetched one round ahead with cp.async into per-warp shared-memory buffers instead of
//     registers (the register file is what limits this kernel to 16 warps per SM), lanes in channel order so that
//     the global and shared sides of every copy are contiguous.
//   * The boustrophedon channel walk permutes the staging SLOT instead of the lane's channel, so a lane streams
//     one channel for the whole segment: constant uv_scale registers, pointer-bump addressing.
//
// This header holds the kernel template and its launcher; standard_grid_window.cu instantiates the 8-wide windows (supports
// 3 / 5 / 7, incl. the fused image + psf and fused imaging-weight modes), standard_grid_window16.cu the 16-wide ones
// (supports 9 / 11 / 13 / 15) -- two translation units so that nvcc compiles them in parallel.
#pragma once
#include "standard_grid.cuh"
#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace cngi {

template <typename T, bool CPLX, int S, int PP, bool DUAL = false, bool IWF = false> struct WinCfg {
    static constexpr int W = (S < 4) ? 4 : (S < 8) ? 8 : 16;     // lanes per item == columns == rows of the register window
    // hysteresis: cells the stamp can move without a slide.  The fit test is one packed compare that needs 2^k - 1, so the
    // spare rows / columns beyond that (supports 11: 5 -> 3) are simply not used.
    static constexpr int SPARE = (W - S >= 7) ? 7 : (W - S >= 3) ? 3 : 1;
    static constexpr int IPW = 32 / W;                           // items per warp
    static constexpr int ITER = W;                               // samples per item per round
    static constexpr int NVC = CPLX ? PP : (PP + 1) / 2;         // accumulator pairs per cell of the grid proper
    static constexpr int NVP = DUAL ? (PP + 1) / 2 : 0;          // fused image + psf pass: (pol 2m, pol 2m+1) pairs of the psf grid
    static constexpr int NV = NVC + NVP;
    static constexpr int TPV = 16 / (int)sizeof(T);              // T's per 16-byte vector
    static constexpr int WD = (NV * 2 + TPV - 1) / TPV * TPV;    // padded weighted-data count per record
#ifndef CNGI_WIN_NS_F32
#define CNGI_WIN_NS_F32 1
#endif
    static constexpr int NS = (sizeof(T) == 4) ? CNGI_WIN_NS_F32 : 1;   // samples consumed per phase-2 iteration
    // staged records are item-major with one pad record per item: rec(i, k) = k * (ITER + 1) + i, so the lanes
    // that stage an item write consecutive records and the items read by one half-warp sit in different banks
    static constexpr int NREC = IPW * (ITER + 1);
    static constexpr int IDX_BYTES = NREC * 16;                  // int4 {packed cell, u row address, v row address, column offset}
    static constexpr int WD_BYTES = NREC * WD * (int)sizeof(T);
    static constexpr int REC_BYTES = IDX_BYTES + WD_BYTES;
    // two raw-sample buffers per warp (cp.async targets): vis, weight, (u, v)
    static constexpr int RAW_BYTES = 32 * (PP * 3 * (int)sizeof(T)) + ITER * 16;
    // fused imaging weights: a two-slot (u, v) ring -- the uv of round r + 1 is in flight while round r is staged, so that
    // the density gather of round r + 1 can be ISSUED one round ahead of its use (it lands in registers: a cp.async gather
    // costs one shared-memory wavefront per lane, measured +64 wavefronts per round on the pipe that co-limits this kernel)
    static constexpr int IW_UV_BYTES = IWF ? ITER * 16 : 0;       // per ring slot
#ifndef CNGI_WIN_SW_SMEM
#define CNGI_WIN_SW_SMEM 0
#endif
    // CNGI_WIN_SW_SMEM: the per-lane sum_weight accumulators (fp64) live in shared memory instead of registers
    static constexpr int SW_BYTES = CNGI_WIN_SW_SMEM ? 32 * PP * 8 * (DUAL ? 2 : 1) : 0;
    static constexpr int WARP_BYTES = REC_BYTES + 2 * RAW_BYTES + 2 * IW_UV_BYTES + SW_BYTES;
    // Tap table.  W <= 8: one copy per rotation, tap[rot][off][W].  W == 16 (supports 9 .. 15): W full rotations would be
    // 105 KB at oversampling 100, so every row is stored TWICE in a row (32 entries) and only the rotations by less than one
    // 16-byte vector are copies (TPV of them): a row rotated by `rot` is the 16 entries that start at entry
    // (16 - (rot - rot % TPV)) % 16 of copy rot % TPV -- still 16-byte aligned vector loads, 53 KB.
    static constexpr bool DOUBLED = W > 8;
    static constexpr int ROW_ELEMS = DOUBLED ? 2 * W : W;
    static constexpr int NROT = DOUBLED ? TPV : W;
    static constexpr int ROW_BYTES = ROW_ELEMS * (int)sizeof(T);
};

struct WinSmem {
    int tap, tapsum, scale, scale_iw, wbuf, total;
};

template <typename Cfg, typename T>
__host__ __device__ inline WinSmem win_smem_layout(int oversampling, int c_n, int warps, bool iw_own_scale = false)
{
    WinSmem L;
    const int n_off = oversampling + 3;
    auto up16 = [](int x) { return (x + 15) / 16 * 16; };
    L.tap = 0;
    L.tapsum = L.tap + up16(Cfg::NROT * n_off * Cfg::ROW_BYTES);
    L.scale = L.tapsum + up16(n_off * (int)sizeof(double));
    L.scale_iw = L.scale + up16(2 * c_n * (int)sizeof(double));
    L.wbuf = L.scale_iw + (iw_own_scale ? up16(2 * c_n * (int)sizeof(double)) : 0);
    L.total = L.wbuf + warps * Cfg::WARP_BYTES;
    return L;
}

// unroll factor of the phase-2 sample loop (development knob: tools/build_variant.sh ... -DCNGI_WIN_UNROLL=2)
#ifndef CNGI_WIN_UNROLL
#define CNGI_WIN_UNROLL 2
#endif
#define CNGI_STR2(x) #x
#define CNGI_STR(x) CNGI_STR2(x)
#define CNGI_WIN_CONSUME_UNROLL _Pragma(CNGI_STR(unroll CNGI_WIN_UNROLL))

#ifndef CNGI_WIN_MINB_F32
#define CNGI_WIN_MINB_F32 4
#endif
#ifndef CNGI_WIN_MINB_F64
#define CNGI_WIN_MINB_F64 3
#endif

#ifndef CNGI_WIN_IW_MINB_F32
#define CNGI_WIN_IW_MINB_F32 4
#endif
#ifndef CNGI_WIN_IW_MINB_F64
#define CNGI_WIN_IW_MINB_F64 3
#endif

#ifndef CNGI_WIN_DUAL_MINB_F32
#define CNGI_WIN_DUAL_MINB_F32 3
#endif
#ifndef CNGI_WIN_DUAL_MINB_F64
#define CNGI_WIN_DUAL_MINB_F64 2
#endif

// calls f(integral_constant<int, j>) for the runtime j in [LO, LO + N) through a binary tree of branches, so that
// the body indexes registers with a compile-time constant
template <int LO, int N, typename F> __device__ __forceinline__ void static_dispatch(int j, F &&f)
{
    if constexpr (N == 1) {
        f(std::integral_constant<int, LO>{});
    } else {
        if (j < LO + N / 2)
            static_dispatch<LO, N / 2>(j, f);
        else
            static_dispatch<LO + N / 2, N / 2>(j, f);
    }
}

template <int LO, int HI, typename F> __device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (LO < HI) {
        f(std::integral_constant<int, LO>{});
        static_for<LO + 1, HI>(f);
    }
}

__device__ __forceinline__ void cp_async_bytes(unsigned dst, const void *src, std::integral_constant<int, 4>)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_bytes(unsigned dst, const void *src, std::integral_constant<int, 8>)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_bytes(unsigned dst, const void *src, std::integral_constant<int, 16>)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// DUAL: one pass grids the image (complex, vis * weight) AND the psf (real, weight) of the same samples -- they share
// every cell index and tap (synthesis_imaging_cube.py:195-211 calls _make_psf and _make_image back to back on the
// same uvw and weights); the psf accumulators ride along as extra (pol 2m, pol 2m+1) pairs of every window cell.
//
// IWF: the weight array holds NATURAL weights and the imaging weight is formed in phase 1 (the weight-degrid pass A4,
// _standard_grid.py:466-518, folded into the gridder): w_img = avg(w) / (f0 * rho[cell] + f1).  The density value is a
// dependent gather (uv -> cell -> rho) that would sit on phase 1's critical path, so it is software-pipelined: the (u, v)
// of round r + 1 is copied (cp.async) while round r is staged, the gather of round r + 1 is issued (cp.async again)
// right after, and lands while the 32 samples of round r are consumed by phase 2.  The imaging weights are never
// written unless the caller asks for them (p.iw_out).
template <typename T, bool CPLX, int S, int PP, int BLK, bool NZ, bool DUAL = false, bool IWF = false>
__global__ void __launch_bounds__(BLK, S > 8 ? (sizeof(T) == 4 ? 3 : 2)   // 16-wide windows: 64 accumulator registers per pol pair
                                        : DUAL ? (sizeof(T) == 4 ? CNGI_WIN_DUAL_MINB_F32 : CNGI_WIN_DUAL_MINB_F64)
                                        : IWF ? (sizeof(T) == 4 ? CNGI_WIN_IW_MINB_F32 : CNGI_WIN_IW_MINB_F64)
                                              : (sizeof(T) == 4 ? CNGI_WIN_MINB_F32 : CNGI_WIN_MINB_F64))
std_grid_window_kernel(StdParams p)
{
    static_assert(!DUAL || CPLX, "the fused image + psf pass grids a complex image");
    using Cfg = WinCfg<T, CPLX, S, PP, DUAL, IWF>;
    constexpr int NVC = Cfg::NVC;
    using CT = typename Cplx<T>::type;
    using P2 = typename Pair<T>::type;
    constexpr int W = Cfg::W, IPW = Cfg::IPW, ITER = Cfg::ITER, NV = Cfg::NV, WD = Cfg::WD, NS = Cfg::NS;
    constexpr int HALF = S / 2;
    constexpr int SPARE = Cfg::SPARE;
    constexpr int ROW_BYTES = Cfg::ROW_BYTES;
    constexpr int kInvalidKey = (int)0x80008000;
    constexpr int kNoWindowKey = 0x7fff7fff;
    constexpr int kFitMask = ~((SPARE << 16) | SPARE);   // SPARE is 2^k - 1
    const unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(16) unsigned char smem[];
    const WinSmem L = win_smem_layout<Cfg, T>(p.oversampling, p.c_n, BLK / 32, IWF && p.iw_own_scale);
    T *tap = reinterpret_cast<T *>(smem + L.tap);
    double *tapsum = reinterpret_cast<double *>(smem + L.tapsum);
    double *scale = reinterpret_cast<double *>(smem + L.scale);
    const int n_off = p.oversampling + 3;
    const int o0 = p.oversampling / 2 + 1;   // table row of oversampling offset 0
    auto tap_of = [&](int q, int off) -> T {  // tap q (0..S-1) of the stamp for oversampling offset `off`
        if (q < 0 || q >= S) return (T)0;
        const int k = abs(p.oversampling * (q - HALF) + off);
        return k < p.table_len ? (T)p.cgk[k] : (T)0;
    };
    // tap[rot][o][j]: slot j holds the tap of the cell that is q = (j - rot) mod W above the lowest stamp cell
    for (int i = threadIdx.x; i < n_off * W; i += BLK) {
        const int o = i / W, q = i % W;
        const T val = tap_of(q, o - o0);
#pragma unroll
        for (int rot = 0; rot < Cfg::NROT; ++rot) {
            const int at = (rot * n_off + o) * Cfg::ROW_ELEMS + ((q + rot) & (W - 1));
            tap[at] = val;
            if (Cfg::DOUBLED) tap[at + W] = val;
        }
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
    const double *scale_iw = scale;   // uv scale of the density grid: the gridder's own unless the geometries differ
    if constexpr (IWF) {
        if (p.iw_own_scale) {
            double *tbl = reinterpret_cast<double *>(smem + L.scale_iw);
            for (int i = threadIdx.x; i < p.c_n; i += BLK) {
                const double f = p.freq[p.c_lo + i];
                tbl[i] = uv_scale_of(f, p.iw_dl, p.iw_n_u);
                tbl[p.c_n + i] = uv_scale_of(f, p.iw_dm, p.iw_n_v);
            }
            scale_iw = tbl;
        }
    }
    __syncthreads();   // the only block-wide barrier

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // per-warp shared memory: staged records (32 x int4 + 32 x WD data) and two raw-sample buffers filled by cp.async
    const unsigned wbuf_s = (unsigned)__cvta_generic_to_shared(smem + L.wbuf + warp * Cfg::WARP_BYTES);
    const unsigned idx_s = wbuf_s;
    const unsigned wd_s = wbuf_s + Cfg::IDX_BYTES;
    const unsigned raw_s = wbuf_s + Cfg::REC_BYTES;
    const unsigned uvr_s = raw_s + 2 * Cfg::RAW_BYTES;            // IWF: (u, v) ring, two slots
    const unsigned tap_s = (unsigned)__cvta_generic_to_shared(tap);
    const int rot_stride = n_off * ROW_BYTES;   // bytes between two rotations of the tap table
    const int G = p.G;
    const int spr = ITER >> p.log2G;   // time steps per round
    const int c_end = p.c_lo + p.c_n;

    // lane roles.  phase 1: lane <-> one channel of one item, `row1`-th time step of the round;
    //              phase 2: lane <-> (item, u residue mod W)
    // (lane = (row1 * IPW + k1) * G + g1: channel-major, so a warp's global loads are contiguous per time step)
    const int g1 = lane & (G - 1);
    const int k1 = (lane >> p.log2G) % IPW;
    const int row1 = (lane >> p.log2G) / IPW;
    const bool uv_lane = (k1 == 0) && (g1 == 0);   // fetches the (u, v) of time step `row1` for the whole warp
    const int k2 = lane / W;
    const int r2 = lane & (W - 1);
    // The channels of an item are consumed boustrophedon (forward on even time steps, backward on odd ones) when
    // they share one image plane, so that consecutive samples of an item are always uv neighbours: the lane keeps
    // its channel and stages into the mirrored slot on odd time steps.
    const bool zigzag = (p.chan_mode == CNGI_CHAN_CONTINUUM) && G > 1;
    const int slot_fwd = k1 * (ITER + 1) + row1 * G + g1;
    const int slot_bwd = zigzag ? (k1 * (ITER + 1) + row1 * G + (G - 1 - g1)) : slot_fwd;

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
    auto lds_one = [](unsigned addr) -> T {
        if constexpr (sizeof(T) == 4) {
            float x;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr));
            return (T)x;
        } else {
            double x;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(addr));
            return (T)x;
        }
    };
    auto lds_idx = [](unsigned addr) -> int4 {
        int4 x;
        asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(addr));
        return x;
    };
    auto lds_f64x2 = [](unsigned addr) -> double2 {
        double2 x;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x.x), "=d"(x.y) : "r"(addr));
        return x;
    };
    auto red_pair = [](auto is_cplx, T *base, int cell, P2 v, T *base1) {   // one accumulator pair into the grid
        // NZ: skip cells that only ever received zero taps (the spare row / column) -- pays when the kernel is bound by
        // the reductions (single-channel tracks on a grid far larger than L2), costs issue slots otherwise
        if constexpr (decltype(is_cplx)::value) {
            if (!NZ || v.x != (T)0 || v.y != (T)0) {
                CT val;
                val.x = v.x, val.y = v.y;
                red_add((CT *)base + cell, val);
            }
        } else {   // pair = (pol 2n, pol 2n+1) of a real grid
            if (!NZ || v.x != (T)0) red_add(base + cell, v.x);
            if (!NZ || v.y != (T)0) red_add(base1 + cell, v.y);
        }
    };

    // One work item per warp by default (the hardware block scheduler balances items of unequal cost: long baselines
    // slide more); with CNGI_WIN_PERSIST=1 the grid is the resident blocks and warps stride over the items.
    for (long long task = (long long)blockIdx.x * (BLK / 32) + warp; task < p.n_tasks;
         task += (long long)gridDim.x * (BLK / 32)) {
        // ---- task decode: (time segment, baseline, pol group, channel span), channel span fastest ------
        const int cspan = (int)(task % p.n_cspan);
        long long rest = task / p.n_cspan;
        const int pgrp = (int)(rest % p.n_pgrp);
        rest /= p.n_pgrp;
        const int b = (int)(rest % p.n_baseline);
        const int seg = (int)(rest / p.n_baseline);
        const int t_lo = seg * p.seg_len;
        const int t_hi = min(p.n_time, t_lo + p.seg_len);
        const int c_base = p.c_lo + cspan * IPW * G;
        const int p0 = pgrp * PP;
        const int npol = min(PP, p.n_pol - p0);

        const int c1 = c_base + k1 * G + g1;
        const bool chan_ok = c1 < c_end;
        const int a_chan1 = chan_ok ? chan_of(p, c1) : 0;
        const int sc1 = chan_ok ? c1 - p.c_lo : 0;   // row of the uv_scale table
#if CNGI_WIN_SW_SMEM
        double *sw_acc = reinterpret_cast<double *>(smem + L.wbuf + warp * Cfg::WARP_BYTES + Cfg::REC_BYTES + 2 * Cfg::RAW_BYTES +
                                                    2 * Cfg::IW_UV_BYTES) + lane * PP;
#else
        double sw_acc[PP];
#endif
#pragma unroll
        for (int ip = 0; ip < PP; ++ip) sw_acc[ip] = 0.0;

        int apol[PP];
#pragma unroll
        for (int ip = 0; ip < PP; ++ip) apol[ip] = (ip < npol) ? pol_of(p, p0 + ip) : 0;
        double bf0[PP], bf1[PP];   // IWF: Briggs factors of this lane's (imaging channel, pol) planes
        bool iw_ok = false;        // IWF: the sample's cell of the density grid exists (set when its gather is issued)
        double rho_reg[PP];        // IWF: the gathered density values of the NEXT sample to be staged (in flight during phase 2)
#pragma unroll
        for (int ip = 0; ip < PP; ++ip) {
            bf0[ip] = bf1[ip] = 0.0;
            rho_reg[ip] = 0.0;
            if constexpr (IWF) {
                const int q = a_chan1 * p.n_ip + apol[ip];
                bf0[ip] = p.iw_bf[q];
                bf1[ip] = p.iw_bf[p.n_ic * p.n_ip + q];
            }
        }
        const int c_item = c_base + k2 * G;   // every channel of an item maps to one image plane (G > 1 only in continuum)
        const int plane2 = (c_item < c_end) ? chan_of(p, c_item) : 0;
        // plane (plane2, apol[ip]) of the grid as T elements; a missing pol (odd pol count) aliases the first one: its
        // accumulators stay zero, so reducing them there is harmless and the flush code needs no pol-count test
        T *gplane[PP];
#pragma unroll
        for (int ip = 0; ip < PP; ++ip)
            gplane[ip] = (T *)p.grid +
                         ((long long)plane2 * p.n_ip + apol[ip < npol ? ip : 0]) * ((long long)p.n_u * p.n_v) * (CPLX ? 2 : 1);
        T *pplane[PP];   // psf planes of the fused pass (real)
#if CNGI_WIN_SW_SMEM
        double *psw_acc = sw_acc + (DUAL ? 32 * PP - 0 : 0);   // second half of the area (other lanes' slots lie in between)
#else
        double psw_acc[PP];
#endif
#pragma unroll
        for (int ip = 0; ip < PP; ++ip) {
            if (DUAL || !CNGI_WIN_SW_SMEM) psw_acc[ip] = 0.0;
            pplane[ip] = nullptr;
            if constexpr (DUAL)
                pplane[ip] = (T *)p.psf_grid + ((long long)plane2 * p.n_ip + apol[ip < npol ? ip : 0]) * ((long long)p.n_u * p.n_v);
        }
        P2 acc[W][NV];
#pragma unroll
        for (int j = 0; j < W; ++j)
#pragma unroll
            for (int n = 0; n < NV; ++n) acc[j][n].x = acc[j][n].y = (T)0;
        // Register window, always inside the grid: grid rows [lo_a, lo_a + W) along v -- the LANE axis, lane r owns the
        // row v == r (mod W) -- and columns [lo_b, lo_b + W) along u -- the ACCUMULATOR axis, accumulator j is the column
        // u == j (mod W).  v is the contiguous axis of the grid, so when a column leaves the window the W lanes of the item
        // reduce W consecutive cells (one coalesced 64-byte request per pol instead of W strided sectors): the reductions
        // are what bounds cube gridding on grids far larger than L2 (~50 G reduction sectors/s measured at L2).
        // wkey = lo_a<<16 | lo_b.
        int lo_a = 0, lo_b = 0, wkey = kNoWindowKey;

        auto my_line = [&]() { return lo_a + ((r2 - lo_a) & (W - 1)); };   // the v in the window with v == r2 (mod W)
        // Reductions only READ the accumulators; they are cleared afterwards with selects outside any divergent
        // region (writes under divergence make ptxas copy the whole accumulator file around the branch).
        auto red_lane = [&]() {   // all W cells of this lane's grid row; accumulator j is the column u == j (mod W)
            const int m = lo_b & (W - 1);
            const int cell0 = (lo_b - m) * p.n_v + my_line();
            static_for<0, W>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                const int cell = cell0 + (j + ((j < m) ? W : 0)) * p.n_v;
                if constexpr (!DUAL) {   // (kept as a plain unrolled loop: the static_for form below costs the product kernel 10 %)
#pragma unroll
                    for (int n = 0; n < NV; ++n)
                        red_pair(std::bool_constant<CPLX>{}, gplane[CPLX ? n : 2 * n], cell, acc[j][n],
                                 gplane[CPLX ? n : (2 * n + 1 < PP ? 2 * n + 1 : 0)]);
                } else {
                    static_for<0, NV>([&](auto nc) {
                        constexpr int n = decltype(nc)::value;
                        if constexpr (n < NVC) {
                            red_pair(std::bool_constant<CPLX>{}, gplane[CPLX ? n : 2 * n], cell, acc[j][n],
                                     gplane[CPLX ? n : (2 * n + 1 < PP ? 2 * n + 1 : 0)]);
                        } else {   // psf pair of the fused pass
                            constexpr int m = n - NVC;
                            red_pair(std::false_type{}, pplane[2 * m], cell, acc[j][n], pplane[2 * m + 1 < PP ? 2 * m + 1 : 0]);
                        }
                    });
                }
            });
        };
        auto red_column = [&](int u) {   // column u of the window: every lane reduces its cell of it (consecutive v)
            const int cell = u * p.n_v + my_line();
            static_dispatch<0, W>(u & (W - 1), [&](auto jc) {
                constexpr int j = decltype(jc)::value;
                if constexpr (!DUAL) {   // (kept as a plain unrolled loop: the static_for form below costs the product kernel 10 %)
#pragma unroll
                    for (int n = 0; n < NV; ++n)
                        red_pair(std::bool_constant<CPLX>{}, gplane[CPLX ? n : 2 * n], cell, acc[j][n],
                                 gplane[CPLX ? n : (2 * n + 1 < PP ? 2 * n + 1 : 0)]);
                } else {
                    static_for<0, NV>([&](auto nc) {
                        constexpr int n = decltype(nc)::value;
                        if constexpr (n < NVC) {
                            red_pair(std::bool_constant<CPLX>{}, gplane[CPLX ? n : 2 * n], cell, acc[j][n],
                                     gplane[CPLX ? n : (2 * n + 1 < PP ? 2 * n + 1 : 0)]);
                        } else {   // psf pair of the fused pass
                            constexpr int m = n - NVC;
                            red_pair(std::false_type{}, pplane[2 * m], cell, acc[j][n], pplane[2 * m + 1 < PP ? 2 * m + 1 : 0]);
                        }
                    });
                }
            });
        };
        // make the stamp whose lowest cell is (v, u) = (need_a, need_b) fit the window, sliding it by the least amount
        auto slide = [&](int need_a, int need_b) {
            int new_a = need_a, new_b = need_b;
            if (wkey != kNoWindowKey) {
                const int da = need_a - lo_a, db = need_b - lo_b;
                new_a = da < 0 ? need_a : (da > SPARE ? need_a - SPARE : lo_a);
                new_b = db < 0 ? need_b : (db > SPARE ? need_b - SPARE : lo_b);
            }
            new_a = min(new_a, p.n_v - W);   // keep the spare rows / columns inside the grid
            new_b = min(new_b, p.n_u - W);
            if (wkey != kNoWindowKey) {
                const int a = my_line();
                const int sh = new_b - lo_b;
                const bool whole = (a < new_a) || (a >= new_a + W) || (sh >= W) || (sh <= -W);   // my grid row leaves
                // columns at window positions [lower, upper) leave
                const int lower = whole ? 0 : (sh > 0 ? 0 : W + sh);
                const int upper = whole ? W : (sh > 0 ? sh : W);
                if (whole) {
                    red_lane();
                } else {
                    for (int pos = lower; pos < upper; ++pos) red_column(lo_b + pos);
                }
                // bit j of `gone_mask`: accumulator j (column u == j mod W) leaves -- the run of window positions [lower, upper)
                // rotated by the window origin; one test per accumulator instead of a subtract / mask / compare each
                const int m = lo_b & (W - 1);
                const unsigned run = ((1u << (upper - lower)) - 1u) << lower;
                const unsigned gone_mask = (run << m) | (run >> (W - m));
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    const bool gone = (gone_mask >> j) & 1u;
#pragma unroll
                    for (int n = 0; n < NV; ++n) {
                        acc[j][n].x = gone ? (T)0 : acc[j][n].x;
                        acc[j][n].y = gone ? (T)0 : acc[j][n].y;
                    }
                }
            }
            lo_a = new_a, lo_b = new_b;
            wkey = (new_a << 16) | new_b;
        };

        // ---- raw samples: cp.async into the warp's shared-memory buffers, one round ahead -----------------
        constexpr int VIS_RAW = 32 * PP * (int)sizeof(CT), W_RAW = 32 * PP * (int)sizeof(T), UV_RAW = ITER * 16;
        constexpr int RAW_BUF = VIS_RAW + W_RAW + UV_RAW;
        const long long s_step = (long long)spr * p.n_baseline * p.n_chan * p.n_pol;
        long long s_next = (((long long)(t_lo + row1) * p.n_baseline + b) * p.n_chan + c1) * p.n_pol + p0;
        const double *uvw_next = p.uvw + ((long long)(t_lo + row1) * p.n_baseline + b) * 3;
        const long long uvw_step = (long long)spr * p.n_baseline * 3;
        const bool vec2 = (PP == 2) && npol == 2 && (p.n_pol & 1) == 0;
        unsigned raw_flag = 0;   // flag bytes of the prefetched sample (2-byte loads cannot go through cp.async)
        auto load_uv = [&](int t0, int slot, const double *src) {   // IWF: (u, v) of the round starting at t0 into a ring slot
            if (uv_lane && chan_ok && (t0 + row1 < t_hi)) {
                cp_async_bytes(uvr_s + slot * Cfg::IW_UV_BYTES + row1 * 16, src, std::integral_constant<int, 8>{});
                cp_async_bytes(uvr_s + slot * Cfg::IW_UV_BYTES + row1 * 16 + 8, src + 1, std::integral_constant<int, 8>{});
            }
        };
        auto load_raw = [&](int t0, int buf) {
            raw_flag = 0;
            if constexpr (IWF) {
                // (u, v) of the round after this one; then the density values of THIS round's samples, whose (u, v) landed
                // before the previous wait (ring slot `buf`): cell of the density grid -> one 8-byte gather per pol
                load_uv(t0 + spr, buf ^ 1, uvw_next + uvw_step);
                iw_ok = false;
                if (chan_ok && (t0 + row1 < t_hi)) {
                    const double2 uv = lds_f64x2(uvr_s + buf * Cfg::IW_UV_BYTES + row1 * 16);
                    CellPos cq;
                    if (locate_centre(uv.x, uv.y, scale_iw[sc1], scale_iw[p.c_n + sc1], p.iw_n_u, p.iw_n_v, cq) &&
                        stamp_inside(cq.uc, cq.vc, 0, p.iw_n_u, p.iw_n_v)) {
                        iw_ok = true;
                        const double *src = p.iw_density + cq.uc * p.iw_ds_u + cq.vc * p.iw_ds_v + a_chan1 * p.iw_ds_c;
                        rho_reg[0] = __ldg(src + apol[0] * p.iw_ds_p);
                        if (PP > 1 && !p.iw_pol_shared && npol > 1) rho_reg[PP - 1] = __ldg(src + apol[PP - 1] * p.iw_ds_p);
                    }
                }
            }
            if (chan_ok && (t0 + row1 < t_hi)) {
                const unsigned base = raw_s + buf * RAW_BUF;
                if (!IWF && uv_lane) {   // one (u, v) per time step of the round: one lane copies it, the row's lanes all read it
                    cp_async_bytes(base + VIS_RAW + W_RAW + row1 * 16, uvw_next, std::integral_constant<int, 8>{});
                    cp_async_bytes(base + VIS_RAW + W_RAW + row1 * 16 + 8, uvw_next + 1, std::integral_constant<int, 8>{});
                }
                const long long s = s_next;
                const unsigned wdst = base + VIS_RAW + lane * (PP * (int)sizeof(T));
                const unsigned vdst = base + lane * (PP * (int)sizeof(CT));
                if (vec2) {   // 2 pols, aligned: one wide copy each
                    cp_async_bytes(wdst, (const T *)p.weight + s, std::integral_constant<int, 2 * (int)sizeof(T)>{});
                    if (!p.do_psf) {
                        cp_async_bytes(vdst, (const CT *)p.vis + s, std::integral_constant<int, 16>{});
                        if (sizeof(CT) == 16) cp_async_bytes(vdst + 16, (const CT *)p.vis + s + 1, std::integral_constant<int, 16>{});
                        if (p.flag) raw_flag = *reinterpret_cast<const unsigned short *>(p.flag + s);
                    }
                } else {
#pragma unroll
                    for (int ip = 0; ip < PP; ++ip) {
                        if (ip < npol) {
                            cp_async_bytes(wdst + ip * (int)sizeof(T), (const T *)p.weight + s + ip,
                                           std::integral_constant<int, (int)sizeof(T)>{});
                            if (!p.do_psf) {
                                cp_async_bytes(vdst + ip * (int)sizeof(CT), (const CT *)p.vis + s + ip,
                                               std::integral_constant<int, (int)sizeof(CT)>{});
                                if (p.flag) raw_flag |= (unsigned)p.flag[s + ip] << (8 * ip);
                            }
                        }
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            s_next += s_step;
            uvw_next += uvw_step;
        };

        // ---- phase 1: locate, mask, stage ----------------------------------------------------------------
        auto stage = [&](int t0, int buf) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();   // (u, v) was copied by the row's uv_lane
            int4 idx = make_int4(kInvalidKey, (int)tap_s, (int)tap_s, 0);
            const int slot = ((t0 + row1) & 1) ? slot_bwd : slot_fwd;
            T wd[WD];
#pragma unroll
            for (int i = 0; i < WD; ++i) wd[i] = (T)0;
            CellPos cp;
            bool ok = chan_ok && (t0 + row1 < t_hi);
            const unsigned base = raw_s + buf * RAW_BUF;
            const T *wsrc = reinterpret_cast<const T *>(smem + (base - (unsigned)__cvta_generic_to_shared(smem)) + VIS_RAW) + lane * PP;
            T raw_w[PP];
#pragma unroll
            for (int ip = 0; ip < PP; ++ip) raw_w[ip] = (T)0;
            if (ok) {
                const double2 uv = lds_f64x2(IWF ? uvr_s + buf * Cfg::IW_UV_BYTES + row1 * 16 : base + VIS_RAW + W_RAW + row1 * 16);
                if constexpr (IWF) {
                    // imaging weights of the sample, operation for operation what iw_degrid_kernel (A4) computes:
                    // off the density grid or NaN uv -> 0; else avg of the two pols (n_pol == 2) or the natural weight,
                    // divided by f0 * rho + f1 where the natural weight and rho are both finite and non-zero
                    auto quotient = [](double num, double den) -> double {
                        return sizeof(T) == 4 ? (double)__fdiv_rn((float)num, (float)den) : __ddiv_rn(num, den);
                    };
                    if (PP == 2 && p.iw_pol_shared && p.n_pol == 2) {
                        // both pols see the same density value and Briggs factors (pol-averaged weights): one division
                        const double w0 = (double)wsrc[0], w1 = (double)wsrc[PP - 1];
                        const double avg = __dmul_rn(__dadd_rn(w0, w1), 0.5);
                        double q = avg;
                        const double r = rho_reg[0];
                        if (iw_ok && !isnan(r) && r != 0.0) q = quotient(avg, __dadd_rn(__dmul_rn(bf0[0], r), bf1[0]));
                        raw_w[0] = (T)(iw_ok ? ((!isnan(w0) && w0 != 0.0) ? q : avg) : 0.0);
                        raw_w[PP - 1] = (T)(iw_ok ? ((!isnan(w1) && w1 != 0.0) ? q : avg) : 0.0);
                    } else {
                        const double avg = (p.n_pol == 2) ? __dmul_rn(__dadd_rn((double)wsrc[0], (double)wsrc[PP - 1]), 0.5) : 0.0;
#pragma unroll
                        for (int ip = 0; ip < PP; ++ip) {
                            if (ip < npol) {
                                double iw = 0.0;
                                if (iw_ok) {
                                    const double w = (double)wsrc[ip];
                                    iw = (p.n_pol == 2) ? avg : w;
                                    const double r = p.iw_pol_shared ? rho_reg[0] : rho_reg[ip];
                                    if (!isnan(w) && w != 0.0 && !isnan(r) && r != 0.0)
                                        iw = quotient(iw, __dadd_rn(__dmul_rn(bf0[ip], r), bf1[ip]));
                                }
                                raw_w[ip] = (T)iw;
                            }
                        }
                    }
                    if (p.iw_out) {   // the caller wants IMAGING_WEIGHT too (s_next already points at the next round)
                        T *dst = (T *)p.iw_out + (s_next - s_step);
#pragma unroll
                        for (int ip = 0; ip < PP; ++ip)
                            if (ip < npol) dst[ip] = raw_w[ip];
                    }
                }
                ok = locate_centre(uv.x, uv.y, scale[sc1], scale[p.c_n + sc1], p.n_u, p.n_v, cp);
            }
            if (ok) ok = stamp_inside(cp.uc, cp.vc, HALF, p.n_u, p.n_v);
            if (ok) {
                CT raw_vis[PP];
                {
                    const CT *vsrc = reinterpret_cast<const CT *>(smem + (base - (unsigned)__cvta_generic_to_shared(smem))) + lane * PP;
#pragma unroll
                    for (int ip = 0; ip < PP; ++ip) {
                        if (!IWF) raw_w[ip] = wsrc[ip];
                        if (!p.do_psf) raw_vis[ip] = vsrc[ip];
                    }
                }
                double wsel[PP], psel[PP];
                bool any = false;
#pragma unroll
                for (int ip = 0; ip < PP; ++ip) {
                    wsel[ip] = 0.0;
                    psel[ip] = 0.0;
                    if (ip < npol) {
                        const T w = raw_w[ip];
                        if constexpr (DUAL) {   // psf mask: the weight alone (_standard_grid.py:327-333,340)
                            if (!(isnan(w) || w == (T)0)) {
                                any = true;
                                psel[ip] = (double)w;
                                wd[2 * NVC + ip] = w;
                            }
                        }
                        T wre = w, wim = (T)0;
                        bool use;
                        if (p.do_psf) {
                            use = !(isnan(w) || w == (T)0);
                        } else {
                            const T a = raw_vis[ip].x, bq = raw_vis[ip].y;
                            const bool flagged = (raw_flag >> (8 * ip)) & 0xffu;
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
                    if constexpr (DUAL) {
#pragma unroll
                        for (int ip = 0; ip < PP; ++ip) psw_acc[ip] += psel[ip] * norm;
                    }
                    const int need_u = cp.uc - HALF, need_v = cp.vc - HALF;
                    // {lowest stamp cell packed v<<16|u, address of the v tap row rotated into lane order, address of the u tap
                    //  row rotated into accumulator order, unused}
                    auto rotated_row = [&](int need, int off) -> int {   // address of slot 0 of the row rotated by need mod W
                        const int rot = need & (W - 1);
                        if constexpr (Cfg::DOUBLED)   // copy rot % TPV, window of 16 entries starting 16 - (rot - rot % TPV) entries in
                            return (int)tap_s + (rot & (Cfg::TPV - 1)) * rot_stride + off * ROW_BYTES +
                                   ((W - (rot & ~(Cfg::TPV - 1))) & (W - 1)) * (int)sizeof(T);
                        return (int)tap_s + rot * rot_stride + off * ROW_BYTES;
                    };
                    // slot j of a rotated row is the tap of the cell == j (mod W): the u row is read whole (accumulator order),
                    // of the v row a lane reads slot r2 -- its own grid row -- so phase 2 adds only a per-lane constant
                    idx = make_int4((need_v << 16) | need_u, rotated_row(need_v, vo), rotated_row(need_u, uo), 0);
                }
            }
            if (sizeof(T) == 4) {
#pragma unroll
                for (int i = 0; i < WD; i += 4)
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(wd_s + (slot * WD + i) * 4), "f"((float)wd[i]),
                                 "f"((float)wd[i + 1]), "f"((float)wd[i + 2]), "f"((float)wd[i + 3])
                                 : "memory");
            } else {
#pragma unroll
                for (int i = 0; i < WD; i += 2)
                    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(wd_s + (slot * WD + i) * 8), "d"((double)wd[i]),
                                 "d"((double)wd[i + 1])
                                 : "memory");
            }
            asm volatile("st.shared.v4.s32 [%0], {%1,%2,%3,%4};" ::"r"(idx_s + slot * 16), "r"(idx.x), "r"(idx.y), "r"(idx.z),
                         "r"(idx.w)
                         : "memory");
        };

        // ---- phase 2: consume ----------------------------------------------------------------------------
        auto fma_sample = [&](const T *wd, T cu, const T *cv) {
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                P2 w2;
                w2.x = wd[2 * n], w2.y = wd[2 * n + 1];
                const P2 t = pk_mul(w2, cu);
#pragma unroll
                for (int j = 0; j < W; ++j) pk_fma_acc(acc[j][n], t, cv[j]);
            }
        };
        // the item's records: running addresses instead of index arithmetic in the loop
        constexpr int WDB = WD * (int)sizeof(T);
        const unsigned rec0 = idx_s + k2 * (ITER + 1) * 16, wrec0 = wd_s + k2 * (ITER + 1) * WDB;
#ifndef CNGI_WIN_PREFETCH
#define CNGI_WIN_PREFETCH 1
#endif
        // Software pipelining of phase 2 (fp32, one sample per iteration): the NEXT record is loaded before this sample's FMAs,
        // with the loop unrolled by two so that the rotating registers are renamed instead of moved.  Measured on C2
        // (continuum / cube, ms): plain 1.660 / 2.874, unrolled by two 1.641 / 2.870, this 1.617 / 2.849; without the unroll
        // 1.76 (seven moves per sample), unrolled by four 1.64 / 3.73 (code size); also fetching the next sample's TAPS one
        // sample ahead (a two-deep pipeline) 1.81 / 3.06.  After the item's last sample the load hits its pad record.
        constexpr bool PREFETCH = CNGI_WIN_PREFETCH && NS == 1 && sizeof(T) == 4;
        auto load_record = [&](unsigned rec, unsigned wrec, int4 &idx, T *wd) {
            idx = lds_idx(rec);
#pragma unroll
            for (int q = 0; q < WD; q += Cfg::TPV)   // (equal strides: the data record sits at a constant distance)
                lds_vec((WDB == 16 ? rec + Cfg::IDX_BYTES : wrec) + q * (int)sizeof(T), wd + q);
        };
        auto consume = [&]() {
            unsigned wrec = wrec0;
            int4 idx_n;
            T wd_n[WD];
            if constexpr (PREFETCH) load_record(rec0, wrec0, idx_n, wd_n);
            CNGI_WIN_CONSUME_UNROLL
            for (unsigned rec = rec0; rec != rec0 + ITER * 16; rec += NS * 16, wrec += NS * WDB) {
                int4 idx[NS];
                T wd[NS][WD], cu[NS], cv[NS][W];
                if constexpr (PREFETCH) {
                    idx[0] = idx_n;
#pragma unroll
                    for (int q = 0; q < WD; ++q) wd[0][q] = wd_n[q];
                } else {
#pragma unroll
                    for (int s = 0; s < NS; ++s) load_record(rec + s * 16, wrec + s * WDB, idx[s], wd[s]);
                }
                int misfit = 0;
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    // neither tap address depends on the window position: the table rows are pre-rotated
                    cu[s] = lds_one(idx[s].y + r2 * (int)sizeof(T));
#pragma unroll
                    for (int q = 0; q < W; q += Cfg::TPV) lds_vec(idx[s].z + q * (int)sizeof(T), cv[s] + q);
                    misfit |= idx[s].x - wkey;   // both 16-bit halves of the difference must be in [0, SPARE]
                }
                if constexpr (PREFETCH) load_record(rec + 16, wrec + WDB, idx_n, wd_n);
                if ((misfit & kFitMask) == 0) {
#pragma unroll
                    for (int s = 0; s < NS; ++s) fma_sample(wd[s], cu[s], cv[s]);
                } else {
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        if (idx[s].x != kInvalidKey) {
                            if ((idx[s].x - wkey) & kFitMask) slide(idx[s].x >> 16, idx[s].x & 0xffff);
                            fma_sample(wd[s], cu[s], cv[s]);
                        }
                    }
                }
            }
        };

        // ---- main loop over rounds ---------------------------------------------------------------------
        int buf = 0;
        if constexpr (IWF) {   // the first round's (u, v) must have landed before its density gather can be issued
            load_uv(t_lo, 0, uvw_next);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
        }
        load_raw(t_lo, 0);
        for (int t0 = t_lo; t0 < t_hi; t0 += spr) {
            stage(t0, buf);
            __syncwarp();
            buf ^= 1;
            load_raw(t0 + spr, buf);   // issues nothing past t_hi
            consume();
            __syncwarp();
        }
        if (wkey != kNoWindowKey) red_lane();

        // ---- sum_weight: lanes that share a channel reduce first, then one reduction per image plane -------
        const int span = IPW * G;   // lanes L and L + span handle the same channel
#pragma unroll
        for (int ip = 0; ip < PP; ++ip) {
            double v = sw_acc[ip];
            for (int o = span; o < 32; o <<= 1) v += __shfl_xor_sync(FULL, v, o);
            const bool lead = (lane < span) && chan_ok && (ip < npol);
            warp_grouped_add(p.sum_weight, a_chan1 * p.n_ip + apol[ip], v, lead);
            if constexpr (DUAL) {
                double q = psw_acc[ip];
                for (int o = span; o < 32; o <<= 1) q += __shfl_xor_sync(FULL, q, o);
                warp_grouped_add(p.psf_sum_weight, a_chan1 * p.n_ip + apol[ip], q, lead);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
//  host launcher
// ------------------------------------------------------------------------------------------------
static inline int env_knob(const char *name, int dflt)   // development knobs (tools/probe_std_grid.py sweeps)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <typename T, bool CPLX, int S, int PP, bool NZ, bool DUAL = false, bool IWF = false>
static int launch_window_t(StdParams p, const cngi_std_grid_args *a, cudaStream_t st)
{
    using Cfg = WinCfg<T, CPLX, S, PP, DUAL, IWF>;
#ifndef CNGI_WIN_BLK
#define CNGI_WIN_BLK 128
#endif
    constexpr int BLK = CNGI_WIN_BLK;
    if (p.n_time == 0 || p.n_baseline == 0 || p.n_chan == 0 || p.n_pol == 0) return CNGI_OK;
    constexpr int kMaxChanWindow = 2048;   // 32 KB of uv-scale table per block at most
    auto kern = std_grid_window_kernel<T, CPLX, S, PP, BLK, NZ, DUAL, IWF>;
    static const int persist = env_knob("CNGI_WIN_PERSIST", 0);
    for (int c_lo = 0; c_lo < p.n_chan; c_lo += kMaxChanWindow) {
        p.c_lo = c_lo;
        p.c_n = std::min(kMaxChanWindow, p.n_chan - c_lo);
        // channels walked per item: only when neighbouring channels share an image plane (an item owns ONE plane)
        int G = a->chan_group;
        if (G <= 0) G = Cfg::ITER;
        if (p.chan_mode != CNGI_CHAN_CONTINUUM) G = 1;
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
        const size_t smem = (size_t)win_smem_layout<Cfg, T>(p.oversampling, p.c_n, BLK / 32, IWF && p.iw_own_scale).total;
        CNGI_REQUIRE(smem <= 227 * 1024, "standard_grid: tap tables too large for shared memory (%zu bytes)", smem);
        CNGI_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // (Measured, round 2: more resident warps do not help this kernel.  96 registers x 5 blocks -- 68 bytes of spills, with
        // the whole unified array carved out as shared memory so that 5 blocks fit -- 1.96 ms against 1.74 ms; 160-thread
        // blocks x 4 at 96 registers 2.00 ms; 24 warps per SM at 80 registers 2.47 ms.  CNGI_WIN_CARVEOUT = 0..100 sets the
        // carve-out preference for such experiments; the default leaves the driver's choice.)
        static const int carve = env_knob("CNGI_WIN_CARVEOUT", -1);
        if (carve >= 0) CNGI_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        int per_sm = 0;
        CNGI_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLK, smem));
        if (per_sm < 1) per_sm = 1;
        const long long resident_warps = (long long)sm_count() * per_sm * (BLK / 32);
        int seg_len = a->time_segment;
        if (seg_len <= 0) {
            // ~16 work items per resident warp keeps the tail small; segments stay long enough that the final
            // flush (W*W cells per item) is amortised
            const long long target = resident_warps * 16;
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
        long long blocks = ceil_div(p.n_tasks, BLK / 32);
        if (persist) blocks = std::min<long long>(blocks, (long long)sm_count() * per_sm);   // resident warps pull the items
        CNGI_REQUIRE(blocks < (1LL << 31), "standard_grid: too many work items for one launch");
        kern<<<(unsigned)blocks, BLK, smem, st>>>(p);
        CNGI_CUDA_TRY(cudaGetLastError());
    }
    return CNGI_OK;
}

}  // namespace cngi
