// standard_degrid_window.cu -- A7 of SURVEY.md section 8: degridding predict with the register-window scheme of
// standard_grid_window.cu run backwards.  (No reference implementation exists: predict_modelvis_image.py:20-40 is a
// stub; the specification is the adjoint of _standard_grid_jit, _standard_grid.py:242-371.)
//
// A work item walks ONE baseline through a time segment (through G neighbouring channels when they share an image
// plane) and keeps the W x W model-grid cells it is under in REGISTERS: W lanes per item, lane r owns the grid row
// v == r (mod W) (v is the contiguous grid axis), register j holds the column u == j (mod W).  A cell is LOADED when it
// enters the window (a new column is one coalesced 64-byte load per pol for the W lanes) instead of being gathered
// again for every sample: the gather kernel in standard_degrid.cu reads 49 cells x pol per sample through L1/L2 and is
// bound by that traffic.  Per sample a lane forms sum_j cb[j] * G[j] (W packed FMAs per pol), scales by its own v tap,
// and the W partial sums are combined with log2(W) shuffle steps.  Phase 1 (32 lanes, one sample each: fp64 bit-exact
// cell / offset arithmetic, staged {packed cell, tap-row addresses, 1 / tap sum}) and the pre-rotated tap table are
// the gridder's; phase 3 writes the round's 32 results with the phase-1 lane mapping (coalesced).
#include "standard_grid.cuh"
#include <algorithm>
#include <type_traits>

// unroll factor of the phase-2 sample loop (development knob)
#ifndef CNGI_DEGRID_UNROLL
#define CNGI_DEGRID_UNROLL 2
#endif
#define CNGI_DG_STR2(x) #x
#define CNGI_DG_STR(x) CNGI_DG_STR2(x)
#define CNGI_DEGRID_CONSUME_UNROLL _Pragma(CNGI_DG_STR(unroll CNGI_DEGRID_UNROLL))

namespace cngi {

struct DgwParams {
    int n_time, n_baseline, n_chan, n_pol;
    int n_ic, n_ip, n_u, n_v;
    const void *grid;
    const double *uvw;
    const double *freq;
    const int64_t *chan_map;
    const double *cgk;
    void *vis;
    double dl, dm;
    int oversampling, chan_mode, normalize, table_len;
    int G, log2G, seg_len, n_seg, n_cspan, c_lo, c_n;
    long long n_tasks;
};

template <typename T, int S, int NP> struct DgwCfg {
    static constexpr int W = (S < 4) ? 4 : 8;
    static constexpr int SPARE = W - S;
    static constexpr int IPW = 32 / W;
    static constexpr int ITER = W;
    static constexpr int NREC = IPW * (ITER + 1);                 // item-major records with one pad record per item
    static constexpr int TPV = 16 / (int)sizeof(T);
    static constexpr int ROW_BYTES = W * (int)sizeof(T);
    static constexpr int IDX_BYTES = NREC * 16;                   // int4 {packed cell, v tap row, rotated u tap row, row offset}
    static constexpr int FAC_BYTES = (NREC * (int)sizeof(T) + 15) / 16 * 16;   // 1 / tap sum (or 1)
    static constexpr int RES_BYTES = NREC * NP * 2 * (int)sizeof(T);
    static constexpr int UV_BYTES = 2 * ITER * 16;                // two (u, v) buffers filled by cp.async
    static constexpr int WARP_BYTES = IDX_BYTES + FAC_BYTES + RES_BYTES + UV_BYTES;
};

template <typename Cfg, typename T> __host__ __device__ inline int dgw_smem_bytes(int oversampling, int c_n, int warps, int *tsum_off,
                                                                            int *scale_off, int *wbuf_off)
{
    const int n_off = oversampling + 3;
    auto up16 = [](int x) { return (x + 15) / 16 * 16; };
    const int tap = up16(Cfg::W * n_off * Cfg::ROW_BYTES);
    const int tsum = up16(n_off * (int)sizeof(T));
    const int scale = up16(2 * c_n * (int)sizeof(double));
    if (tsum_off) *tsum_off = tap;
    if (scale_off) *scale_off = tap + tsum;
    if (wbuf_off) *wbuf_off = tap + tsum + scale;
    return tap + tsum + scale + warps * Cfg::WARP_BYTES;
}

template <typename T, int S, int NP, int BLK>
__global__ void __launch_bounds__(BLK, (sizeof(T) == 4 ? 4 : 3)) std_degrid_window_kernel(DgwParams p)
{
    using Cfg = DgwCfg<T, S, NP>;
    using CT = typename Cplx<T>::type;
    using P2 = typename Pair<T>::type;
    constexpr int W = Cfg::W, IPW = Cfg::IPW, ITER = Cfg::ITER, SPARE = Cfg::SPARE, ROW_BYTES = Cfg::ROW_BYTES;
    constexpr int HALF = S / 2;
    constexpr int kInvalidKey = (int)0x80008000;
    constexpr int kNoWindowKey = 0x7fff7fff;
    constexpr int kFitMask = ~((SPARE << 16) | SPARE);
    const unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(16) unsigned char smem[];
    int tsum_off, scale_off, wbuf_off;
    dgw_smem_bytes<Cfg, T>(p.oversampling, p.c_n, BLK / 32, &tsum_off, &scale_off, &wbuf_off);
    T *tap = reinterpret_cast<T *>(smem);
    T *tsum = reinterpret_cast<T *>(smem + tsum_off);
    double *scale = reinterpret_cast<double *>(smem + scale_off);
    const int n_off = p.oversampling + 3;
    const int o0 = p.oversampling / 2 + 1;
    auto tap_of = [&](int q, int off) -> T {
        if (q < 0 || q >= S) return (T)0;
        const int k = abs(p.oversampling * (q - HALF) + off);
        return k < p.table_len ? (T)p.cgk[k] : (T)0;
    };
    for (int i = threadIdx.x; i < n_off * W; i += BLK) {
        const int o = i / W, q = i % W;
        const T val = tap_of(q, o - o0);
#pragma unroll
        for (int rot = 0; rot < W; ++rot) tap[(rot * n_off + o) * W + ((q + rot) & (W - 1))] = val;
    }
    for (int i = threadIdx.x; i < n_off; i += BLK) {
        T sum = (T)0;
        for (int q = 0; q < S; ++q) sum += tap_of(q, i - o0);
        tsum[i] = sum;
    }
    for (int i = threadIdx.x; i < p.c_n; i += BLK) {
        const double f = p.freq[p.c_lo + i];
        scale[i] = uv_scale_of(f, p.dl, p.n_u);
        scale[p.c_n + i] = uv_scale_of(f, p.dm, p.n_v);
    }
    __syncthreads();   // the only block-wide barrier

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char *wbuf = smem + wbuf_off + warp * Cfg::WARP_BYTES;
    int4 *idx_arr = reinterpret_cast<int4 *>(wbuf);
    T *fac_arr = reinterpret_cast<T *>(wbuf + Cfg::IDX_BYTES);
    CT *res_arr = reinterpret_cast<CT *>(wbuf + Cfg::IDX_BYTES + Cfg::FAC_BYTES);
    const unsigned uv_s = (unsigned)__cvta_generic_to_shared(wbuf + Cfg::IDX_BYTES + Cfg::FAC_BYTES + Cfg::RES_BYTES);
    const unsigned tap_s = (unsigned)__cvta_generic_to_shared(tap);
    const int rot_stride = n_off * ROW_BYTES;
    const int G = p.G;
    const int spr = ITER >> p.log2G;
    const int c_end = p.c_lo + p.c_n;

    // phase-1 / phase-3 role: lane = (row1 * IPW + k1) * G + g1 (channel-major);  phase-2 role: lane = k2 * W + r2
    const int g1 = lane & (G - 1);
    const int k1 = (lane >> p.log2G) % IPW;
    const int row1 = (lane >> p.log2G) / IPW;
    const bool uv_lane = (k1 == 0) && (g1 == 0);
    const int k2 = lane / W;
    const int r2 = lane & (W - 1);
    const bool zigzag = (p.chan_mode == CNGI_CHAN_CONTINUUM) && G > 1;
    const int slot_fwd = k1 * (ITER + 1) + row1 * G + g1;
    const int slot_bwd = zigzag ? (k1 * (ITER + 1) + row1 * G + (G - 1 - g1)) : slot_fwd;

    auto lds_vec = [](unsigned addr, T *dst) {
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
    auto shfl_pair = [&](P2 v, int o) -> P2 {
        P2 r;
        r.x = __shfl_xor_sync(FULL, v.x, o);
        r.y = __shfl_xor_sync(FULL, v.y, o);
        return r;
    };

    for (long long task = (long long)blockIdx.x * (BLK / 32) + warp; task < p.n_tasks;
         task += (long long)gridDim.x * (BLK / 32)) {
        const int cspan = (int)(task % p.n_cspan);
        long long rest = task / p.n_cspan;
        const int b = (int)(rest % p.n_baseline);
        const int seg = (int)(rest / p.n_baseline);
        const int t_lo = seg * p.seg_len;
        const int t_hi = min(p.n_time, t_lo + p.seg_len);
        const int c_base = p.c_lo + cspan * IPW * G;

        const int c1 = c_base + k1 * G + g1;
        const bool chan_ok = c1 < c_end;
        const int sc1 = chan_ok ? c1 - p.c_lo : 0;
        const int c_item = c_base + k2 * G;
        int plane2 = 0;
        if (c_item < c_end)
            plane2 = p.chan_mode == CNGI_CHAN_CUBE ? c_item : (p.chan_mode == CNGI_CHAN_CONTINUUM ? 0 : (int)p.chan_map[c_item]);
        const CT *gplane[NP];
#pragma unroll
        for (int ip = 0; ip < NP; ++ip)
            gplane[ip] = (const CT *)p.grid + ((long long)plane2 * p.n_ip + ip) * ((long long)p.n_u * p.n_v);
        P2 g[W][NP];   // model-grid cells of the window: register j <-> column u == j (mod W), this lane's row
#pragma unroll
        for (int j = 0; j < W; ++j)
#pragma unroll
            for (int n = 0; n < NP; ++n) g[j][n].x = g[j][n].y = (T)0;
        int lo_a = 0, lo_b = 0, wkey = kNoWindowKey;   // rows [lo_a, lo_a + W) along v, columns [lo_b, lo_b + W) along u

        // make the stamp whose lowest cell is (v, u) = (need_a, need_b) fit the window (least slide, clamped inside the
        // grid) and load the cells that entered it
        auto slide = [&](int need_a, int need_b) {
            const bool first = wkey == kNoWindowKey;
            int new_a = need_a, new_b = need_b;
            if (!first) {
                const int da = need_a - lo_a, db = need_b - lo_b;
                new_a = da < 0 ? need_a : (da > SPARE ? need_a - SPARE : lo_a);
                new_b = db < 0 ? need_b : (db > SPARE ? need_b - SPARE : lo_b);
            }
            new_a = min(new_a, p.n_v - W);
            new_b = min(new_b, p.n_u - W);
            const int a_old = lo_a + ((r2 - lo_a) & (W - 1));
            const int a_new = new_a + ((r2 - new_a) & (W - 1));
            const bool row_changed = first || (a_new != a_old);
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const int u_old = lo_b + ((j - lo_b) & (W - 1));
                const int u_new = new_b + ((j - new_b) & (W - 1));
                if (row_changed || u_new != u_old) {
                    const int cell = u_new * p.n_v + a_new;
#pragma unroll
                    for (int n = 0; n < NP; ++n) {
                        const CT x = gplane[n][cell];
                        g[j][n].x = x.x, g[j][n].y = x.y;
                    }
                }
            }
            lo_a = new_a, lo_b = new_b;
            wkey = (new_a << 16) | new_b;
        };

        // ---- (u, v) of a round: one lane per time step copies it with cp.async one round ahead ----------------
        const double *uvw_next = p.uvw + ((long long)(t_lo + row1) * p.n_baseline + b) * 3;
        const long long uvw_step = (long long)spr * p.n_baseline * 3;
        auto load_uv = [&](int t0, int buf) {
            if (uv_lane && (t0 + row1 < t_hi)) {
                const unsigned dst = uv_s + buf * (ITER * 16) + row1 * 16;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(uvw_next) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8), "l"(uvw_next + 1) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            uvw_next += uvw_step;
        };

        // ---- phase 1: locate, stage ------------------------------------------------------------------------------
        auto stage = [&](int t0, int buf) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            int4 idx = make_int4(kInvalidKey, (int)tap_s, (int)tap_s, 0);
            T fac = (T)1;
            const int slot = ((t0 + row1) & 1) ? slot_bwd : slot_fwd;
            if (chan_ok && (t0 + row1 < t_hi)) {
                double2 uv;
                asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(uv.x), "=d"(uv.y) : "r"(uv_s + buf * (ITER * 16) + row1 * 16));
                CellPos cp;
                bool ok = locate_centre(uv.x, uv.y, scale[sc1], scale[p.c_n + sc1], p.n_u, p.n_v, cp);
                if (ok) ok = stamp_inside(cp.uc, cp.vc, HALF, p.n_u, p.n_v);
                if (ok) {
                    const int uo = oversample_offset(cp.uc, cp.u_pos, p.oversampling) + o0;
                    const int vo = oversample_offset(cp.vc, cp.v_pos, p.oversampling) + o0;
                    const int need_u = cp.uc - HALF, need_v = cp.vc - HALF;
                    // both tap rows pre-rotated: slot j is the tap of the cell == j (mod W), so a lane reads slot r2 of the
                    // v row (its own grid row) with no per-sample address arithmetic, and the u row whole
                    idx = make_int4((need_v << 16) | need_u, (int)tap_s + (need_v & (W - 1)) * rot_stride + vo * ROW_BYTES,
                                    (int)tap_s + (need_u & (W - 1)) * rot_stride + uo * ROW_BYTES, 0);
                    if (p.normalize) fac = (T)1 / (tsum[uo] * tsum[vo]);   // 1 / (sum of the S*S taps)
                }
            }
            idx_arr[slot] = idx;
            fac_arr[slot] = fac;
        };

        // ---- phase 2: consume ----------------------------------------------------------------------------------------
        auto consume = [&]() {
            CNGI_DEGRID_CONSUME_UNROLL
            for (int i = 0; i < ITER; ++i) {
                const int slot = k2 * (ITER + 1) + i;
                const int4 idx = idx_arr[slot];
                const T fac = fac_arr[slot];
                const T ca = lds_one(idx.y + r2 * (int)sizeof(T));
                T cb[W];
#pragma unroll
                for (int q = 0; q < W; q += Cfg::TPV) lds_vec(idx.z + q * (int)sizeof(T), cb + q);
                P2 out[NP];
#pragma unroll
                for (int n = 0; n < NP; ++n) out[n].x = out[n].y = (T)0;
                if (idx.x != kInvalidKey) {
                    if ((idx.x - wkey) & kFitMask) slide(idx.x >> 16, idx.x & 0xffff);
#pragma unroll
                    for (int n = 0; n < NP; ++n) {
                        P2 s0, s1;   // two chains per pol for instruction-level parallelism
                        s0.x = s0.y = s1.x = s1.y = (T)0;
#pragma unroll
                        for (int j = 0; j < W; j += 2) {
                            pk_fma_acc(s0, g[j][n], cb[j]);
                            pk_fma_acc(s1, g[j + 1][n], cb[j + 1]);
                        }
                        s0.x += s1.x, s0.y += s1.y;
                        out[n] = pk_mul(s0, ca);
                    }
                }
#pragma unroll
                for (int o = W / 2; o > 0; o >>= 1) {
#pragma unroll
                    for (int n = 0; n < NP; ++n) {
                        const P2 other = shfl_pair(out[n], o);
                        out[n].x += other.x, out[n].y += other.y;
                    }
                }
                if (r2 == 0) {
#pragma unroll
                    for (int n = 0; n < NP; ++n) {
                        CT v;
                        v.x = out[n].x * fac, v.y = out[n].y * fac;   // samples the gridder would skip yield exactly 0
                        res_arr[slot * NP + n] = v;
                    }
                }
            }
        };

        // ---- main loop over rounds -----------------------------------------------------------------------------------
        long long s_out = (((long long)(t_lo + row1) * p.n_baseline + b) * p.n_chan + c1) * NP;
        const long long s_step = (long long)spr * p.n_baseline * p.n_chan * NP;
        int buf = 0;
        load_uv(t_lo, 0);
        for (int t0 = t_lo; t0 < t_hi; t0 += spr) {
            stage(t0, buf);
            __syncwarp();
            buf ^= 1;
            load_uv(t0 + spr, buf);
            consume();
            __syncwarp();
            if (chan_ok && (t0 + row1 < t_hi)) {   // phase 3: this lane's sample of the round (coalesced along channels)
                const int slot = ((t0 + row1) & 1) ? slot_bwd : slot_fwd;
                CT *dst = (CT *)p.vis + s_out;
#pragma unroll
                for (int n = 0; n < NP; ++n) dst[n] = res_arr[slot * NP + n];
            }
            s_out += s_step;
            __syncwarp();
        }
    }
}

template <typename T, int S, int NP> static int launch_degrid_window_t(DgwParams p, cudaStream_t st)
{
    using Cfg = DgwCfg<T, S, NP>;
    constexpr int BLK = 128;
    constexpr int kMaxChanWindow = 2048;
    auto kern = std_degrid_window_kernel<T, S, NP, BLK>;
    for (int c_lo = 0; c_lo < p.n_chan; c_lo += kMaxChanWindow) {
        p.c_lo = c_lo;
        p.c_n = std::min(kMaxChanWindow, p.n_chan - c_lo);
        int G = (p.chan_mode == CNGI_CHAN_CONTINUUM) ? Cfg::ITER : 1;
        while (G > 1 && (Cfg::IPW * G / 2) >= p.c_n) G >>= 1;
        int log2G = 0;
        while ((1 << (log2G + 1)) <= G) ++log2G;
        G = 1 << log2G;
        p.G = G, p.log2G = log2G;
        const int spr = Cfg::ITER / G;
        p.n_cspan = (int)ceil_div(p.c_n, Cfg::IPW * G);
        const long long per_seg = (long long)p.n_baseline * p.n_cspan;
        const size_t smem = (size_t)dgw_smem_bytes<Cfg, T>(p.oversampling, p.c_n, BLK / 32, nullptr, nullptr, nullptr);
        CNGI_REQUIRE(smem <= 227 * 1024, "standard_degrid: tap tables too large for shared memory (%zu bytes)", smem);
        CNGI_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // ~16 work items per resident warp; a segment start costs one full window load (W*W cells per item)
        const long long target = (long long)sm_count() * 16 * 16;
        long long n_seg = ceil_div(target, per_seg);
        if (n_seg < 1) n_seg = 1;
        int seg_len = (int)ceil_div(p.n_time, n_seg);
        const int min_len = 64 * spr / Cfg::ITER > 8 ? 64 * spr / Cfg::ITER : 8;
        if (seg_len < min_len) seg_len = min_len;
        seg_len = (int)(ceil_div(seg_len, spr) * spr);
        p.seg_len = seg_len;
        p.n_seg = (int)ceil_div(p.n_time, seg_len);
        p.n_tasks = per_seg * p.n_seg;
        const long long blocks = ceil_div(p.n_tasks, BLK / 32);
        CNGI_REQUIRE(blocks < (1LL << 31), "standard_degrid: too many work items for one launch");
        kern<<<(unsigned)blocks, BLK, smem, st>>>(p);
        CNGI_CUDA_TRY(cudaGetLastError());
    }
    return CNGI_OK;
}

bool degrid_window_supported(const cngi_std_degrid_args *a)
{
    if (!(a->support == 3 || a->support == 5 || a->support == 7)) return false;
    if (a->pol_map != nullptr || !(a->n_pol == 1 || a->n_pol == 2) || a->n_pol > a->n_imag_pol) return false;
    const int w = a->support < 4 ? 4 : 8;
    if (a->n_u < w || a->n_v < w || a->n_u > 32767 || a->n_v > 32767 || a->oversampling < 1) return false;
    if ((long long)a->oversampling * (a->support / 2 + 1) > 8192) return false;
    const int tsz = a->precision == CNGI_F32 ? 4 : 8;
    return (long long)(a->oversampling + 3) * (w * w * tsz + tsz) <= 56 * 1024;
}

int launch_degrid_window(const cngi_std_degrid_args *a, cudaStream_t st)
{
    DgwParams p{};
    p.n_time = (int)a->n_time, p.n_baseline = (int)a->n_baseline, p.n_chan = (int)a->n_chan, p.n_pol = (int)a->n_pol;
    p.n_ic = (int)a->n_imag_chan, p.n_ip = (int)a->n_imag_pol, p.n_u = (int)a->n_u, p.n_v = (int)a->n_v;
    p.grid = a->model_grid, p.uvw = a->uvw, p.freq = a->freq_chan, p.chan_map = a->chan_map;
    p.cgk = a->cgk_1D, p.vis = a->vis, p.dl = a->delta_lm[0], p.dm = a->delta_lm[1];
    p.oversampling = a->oversampling, p.chan_mode = a->chan_mode, p.normalize = a->normalize;
    p.table_len = a->oversampling * (a->support / 2 + 1);
#define CNGI_DGW(TT, SS) (a->n_pol == 1 ? launch_degrid_window_t<TT, SS, 1>(p, st) : launch_degrid_window_t<TT, SS, 2>(p, st))
    if (a->precision == CNGI_F32) {
        switch (a->support) {
            case 3: return CNGI_DGW(float, 3);
            case 5: return CNGI_DGW(float, 5);
            default: return CNGI_DGW(float, 7);
        }
    }
    switch (a->support) {
        case 3: return CNGI_DGW(double, 3);
        case 5: return CNGI_DGW(double, 5);
        default: return CNGI_DGW(double, 7);
    }
#undef CNGI_DGW
}

}  // namespace cngi
