// aperture_grid.cu -- A5/A6 of SURVEY.md section 8: A-projection / mosaic gridders.
//   A5  _aperture_grid_jit         (/root/reference/ngcasa/imaging/_imaging_utils/_aperture_grid.py:376-513)
//   A6  _aperture_weight_grid_jit  (_aperture_grid.py:180-291)
//
// The reference multiplies the whole CF stack by the pointing's phase gradient whenever the field changes
// (_aperture_grid.py:428-430, O(n_cf * 160^2) per change); here the product CF[cf] * PG[field] is formed
// on the fly for exactly the taps a sample reads, so the CF stack (real, L2 resident) and the per-field
// phase gradients are only ever read.
//
// A5: the oversampled CF is first re-laid out per call into OFFSET-MAJOR order and pre-multiplied by the phase gradient,
//     taps[field][cf][u offset][v offset][iu][iv] (complex), so the Su x Sv taps a sample needs are one contiguous block
//     instead of a stride-`oversampling` gather through a 160 x 160 array.  The gridding kernel then gives LW lanes to a
//     sample, lane j <-> grid column vc + iv_j (the fastest axis): per stamp row the lanes read one contiguous run of
//     taps and issue one contiguous run of reductions (REDG.F32x2 / 2 x REDG.F64), 4 cells per 32-byte L2 sector.
//     The fp64 cell / offset arithmetic is done once per sample (lane L <-> sample L) and handed out with shuffles.
// A6: every sample stamps the SAME Su x Sv cells at the grid centre, so weights are first summed per
//     (field, cf_baseline, cf_chan, cf_pol, image plane) bucket (warp-aggregated REDG into a small table) and a
//     second tiny kernel multiplies each bucket by its CF*PG taps -- O(n_samples) + O(n_buckets * S^2)
//     instead of O(n_samples * S^2) colliding atomics.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>

namespace cngi {

struct ApParams {
    int n_time, n_baseline, n_chan, n_pol;
    int n_ic, n_ip, n_u, n_v;
    const void *vis;
    const void *weight;
    const uint8_t *flag;
    const double *uvw;
    const double *freq;
    const int64_t *chan_map, *pol_map, *field, *field_id, *cf_b_map, *cf_c_map, *cf_p_map, *support;
    const double *ck;
    const double2 *pg;
    void *grid;
    double *sum_weight;
    double dl, dm;
    int n_field, n_cfb, n_cfc, n_cfp, n_cu, n_cv;
    int os_u, os_v, max_support, do_psf, chan_mode;
    const double *scale;   // [2, n_chan] uv_scale table
    const void *taps;      // A5: offset-major pre-multiplied taps (see aperture_build_taps_kernel)
    int smax, n_off_u, n_off_v;
    const double2 *tapnorm; // track kernel: sum of the taps of each (field, cf, ou, ov) block
    int G, log2G, seg_len, n_seg, n_cspan, n_pgrp;
    long long n_tasks;
    double *buckets;       // A6: [n_field, n_cfb, n_cfc, n_cfp, n_ic, n_ip] weight sums
};

__device__ __forceinline__ int ap_chan_of(const ApParams &p, int c)
{
    if (p.chan_mode == CNGI_CHAN_CUBE) return c;
    if (p.chan_mode == CNGI_CHAN_CONTINUUM) return 0;
    return (int)p.chan_map[c];
}

__device__ __forceinline__ int ap_find_field(const ApParams &p, long long f)
{
    for (int i = 0; i < p.n_field; ++i)
        if (p.field_id[i] == f) return i;
    return -1;
}

// per-sample geometry shared by A5 and A6: row/field skip, centre cell, full-support bounds test (:397,447)
__device__ __forceinline__ bool ap_locate(const ApParams &p, long long tb, int c, int &field_indx, CellPos &cp)
{
    const long long f = p.field[tb];
    if (!(f > -1)) return false;
    field_indx = ap_find_field(p, f);
    if (field_indx < 0) return false;
    if (!locate_centre(p.uvw[tb * 3], p.uvw[tb * 3 + 1], p.scale[c], p.scale[p.n_chan + c], p.n_u, p.n_v, cp)) return false;
    return stamp_inside(cp.uc, cp.vc, p.max_support, p.n_u, p.n_v);
}

// taps[(((field * n_cf + cf) * n_off_u + ou) * n_off_v + ov) * smax + iu_t) * smax + iv_t]
//   = CF[cf][os_u * (iu_t - smax/2) + (ou - os_u/2 - 1) + n_cu/2][os_v * (iv_t - smax/2) + (ov - os_v/2 - 1) + n_cv/2] * PG[field][same]
// i.e. exactly the element _aperture_grid_jit reads for stamp position (iu, iv) at oversampling offset (ou, ov)
// (_aperture_grid.py:448-451,484-499), as (k + 0j) * pg like the reference's conv_kernel * phase_gradient (:429).
template <typename T> __global__ void aperture_build_taps_kernel(ApParams p, typename Cplx<T>::type *taps, long long n_elems)
{
    using CT = typename Cplx<T>::type;
    const int n_cf = p.n_cfb * p.n_cfc * p.n_cfp;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_elems; e += (long long)gridDim.x * blockDim.x) {
        long long r = e;
        const int iv_t = (int)(r % p.smax);
        r /= p.smax;
        const int iu_t = (int)(r % p.smax);
        r /= p.smax;
        const int ov = (int)(r % p.n_off_v);
        r /= p.n_off_v;
        const int ou = (int)(r % p.n_off_u);
        r /= p.n_off_u;
        const int cf = (int)(r % n_cf);
        const int field = (int)(r / n_cf);
        const int cf_u = p.os_u * (iu_t - p.smax / 2) + (ou - p.os_u / 2 - 1) + p.n_cu / 2;
        const int cf_v = p.os_v * (iv_t - p.smax / 2) + (ov - p.os_v / 2 - 1) + p.n_cv / 2;
        CT out;
        out.x = out.y = (T)0;
        if (cf_u >= 0 && cf_u < p.n_cu && cf_v >= 0 && cf_v < p.n_cv) {
            const double k = p.ck[((long long)cf * p.n_cu + cf_u) * p.n_cv + cf_v];
            const double2 g = p.pg[((long long)field * p.n_cu + cf_u) * p.n_cv + cf_v];
            out.x = (T)(k * g.x);
            out.y = (T)(k * g.y);
        }
        taps[e] = out;
    }
}

template <typename T, int LW> __global__ void __launch_bounds__(256) aperture_grid_kernel(ApParams p)
{
    using CT = typename Cplx<T>::type;
    constexpr int SPW = 32 / LW;
    const unsigned FULL = 0xffffffffu;
    const long long total = (long long)p.n_time * p.n_baseline * p.n_chan;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LW, lj = lane % LW;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long plane_cells = (long long)p.n_u * p.n_v;
    const int n_cf = p.n_cfb * p.n_cfc * p.n_cfp;
    const int shalf = p.smax / 2;
    const int iv = lj - shalf;                      // stamp column of this lane
    // running sum_weight of this lane group (one reduction when the image plane changes / at the end)
    double sw_acc = 0.0;
    int sw_slot = -1;

    for (long long base = warp0 * 32; base < total; base += n_warps * 32) {
        // ---- step 1: lane L locates sample base + L ------------------------------------------------------------------
        const long long mine = base + lane;
        int pk_cell = -1, pk_off = 0, pk_cf = 0, pk_plane = 0;
        if (mine < total) {
            const int c = (int)(mine % p.n_chan);
            const long long tb = mine / p.n_chan;
            int field_indx;
            CellPos cp;
            if (ap_locate(p, tb, c, field_indx, cp)) {
                const int ou = oversample_offset(cp.uc, cp.u_pos, p.os_u) + p.os_u / 2 + 1;   // :448-451 (centre folded into the table)
                const int ov = oversample_offset(cp.vc, cp.v_pos, p.os_v) + p.os_v / 2 + 1;
                pk_cell = cp.uc | (cp.vc << 16);
                pk_off = ou | (ov << 16);
                pk_cf = ((int)p.cf_b_map[tb % p.n_baseline] * p.n_cfc + (int)p.cf_c_map[c]) | (field_indx << 20);
                pk_plane = ap_chan_of(p, c);
            }
        }
        // ---- step 2: LW lanes per sample, lane lj <-> stamp column iv ------------------------------------------------------
#pragma unroll 1
        for (int it = 0; it < LW; ++it) {
            const int src = it * SPW + sub;
            const int q_cell = __shfl_sync(FULL, pk_cell, src);
            const int q_off = __shfl_sync(FULL, pk_off, src);
            const int q_cf = __shfl_sync(FULL, pk_cf, src);
            const int a_chan = __shfl_sync(FULL, pk_plane, src);
            const long long idx = base + src;
            const bool ok = q_cell != -1;
            const int uc = q_cell & 0xffff, vc = (int)((unsigned)q_cell >> 16);
            const int ou = q_off & 0xffff, ov = (int)((unsigned)q_off >> 16);
            const int cf_bc = q_cf & 0xfffff, field_indx = q_cf >> 20;
            for (int ip = 0; ip < p.n_pol; ++ip) {   // uniform trip count: shuffles below
                double w = 0.0, wre = 0.0, wim = 0.0;
                bool use = ok;
                if (use) {
                    const long long s = idx * p.n_pol + ip;
                    w = (double)((const T *)p.weight)[s];
                    wre = w;
                    if (!p.do_psf) {
                        const CT d = ((const CT *)p.vis)[s];
                        weighted_vis((double)d.x, (double)d.y, w, wre, wim);
                        if (p.flag && p.flag[s]) wre = nan("");
                    }
                    use = !masked(wre, wim);
                }
                const int a_pol = p.pol_map ? (int)p.pol_map[ip] : ip;
                T nre = (T)0, nim = (T)0;
                if (use) {
                    const int cf = cf_bc * p.n_cfp + (int)p.cf_p_map[ip];
                    const int su = (int)p.support[cf * 2], sv = (int)p.support[cf * 2 + 1];
                    if (iv >= -(sv / 2) && iv < sv - sv / 2) {
                        const CT *tp = (const CT *)p.taps +
                                       ((((long long)field_indx * n_cf + cf) * p.n_off_u + ou) * p.n_off_v + ov) * p.smax * p.smax + lj;
                        CT *col = (CT *)p.grid + ((long long)a_chan * p.n_ip + a_pol) * plane_cells + vc + iv;
                        const T dre = (T)wre, dim = (T)wim;
                        for (int iu = -(su / 2); iu < su - su / 2; ++iu) {
                            const CT t = tp[(iu + shalf) * p.smax];
                            CT val;
                            val.x = t.x * dre - t.y * dim;
                            val.y = t.x * dim + t.y * dre;
                            red_add(col + (long long)(uc + iu) * p.n_v, val);
                            nre += t.x;
                            nim += t.y;
                        }
                    }
                }
#pragma unroll
                for (int o = LW / 2; o > 0; o >>= 1) {
                    nre += __shfl_xor_sync(FULL, nre, o);
                    nim += __shfl_xor_sync(FULL, nim, o);
                }
                if (lj == 0 && use) {   // psf: w * Re(norm); image: w * Re(norm^2)   (:508-511)
                    const double nr = (double)nre, ni = (double)nim;
                    const double term = p.do_psf ? w * nr : w * (nr * nr - ni * ni);
                    const int slot = a_chan * p.n_ip + a_pol;
                    if (slot != sw_slot) {
                        if (sw_slot >= 0 && sw_acc != 0.0) atomicAdd(p.sum_weight + sw_slot, sw_acc);
                        sw_slot = slot;
                        sw_acc = 0.0;
                    }
                    sw_acc += term;
                }
            }
        }
    }
    if (sw_slot >= 0 && sw_acc != 0.0) atomicAdd(p.sum_weight + sw_slot, sw_acc);
}

// ------------------------------------------------------------------------------------------------
//  A5 track kernel: the aperture gridder with the standard gridder's register-window scheme
// ------------------------------------------------------------------------------------------------
// Same decomposition as std_grid_track_kernel (standard_grid.cu): a work item walks one baseline through time (and G
// channels in continuum mode), W lanes per item, lane r owns the grid column u == r (mod W) of a W x W register window
// and keeps W accumulators (rows v == s (mod W)) per polarisation; cells are reduced into the grid (REDG) only when
// they leave the window.  W = 16 covers supports up to 15 with one spare column / row of hysteresis.
// The CF is not separable, so taps are not staged: every lane fetches its W taps per sample straight from the
// block-major tap table tapsW[block][iu][iv] (block = (field, cf, u offset, v offset), zero outside the CF's own
// support), where the W lanes of an item read one contiguous 128-byte row per load.  The per-block tap sum needed by
// sum_weight (_aperture_grid.py:508-511) is tabulated once per call (tapnorm).
template <typename T, int W> __global__ void aperture_build_blocks_kernel(ApParams p, typename Cplx<T>::type *tapsW, double2 *tapnorm)
{
    using CT = typename Cplx<T>::type;
    const int n_cf = p.n_cfb * p.n_cfc * p.n_cfp;
    long long r = blockIdx.x;                       // one CUDA block per tap block
    const int ov = (int)(r % p.n_off_v);
    r /= p.n_off_v;
    const int ou = (int)(r % p.n_off_u);
    r /= p.n_off_u;
    const int cf = (int)(r % n_cf);
    const int field = (int)(r / n_cf);
    const int su = (int)p.support[cf * 2], sv = (int)p.support[cf * 2 + 1];
    const int shalf = p.smax / 2;
    double sre = 0.0, sim = 0.0;
    for (int e = threadIdx.x; e < W * W; e += blockDim.x) {
        const int iu_t = e / W, iv_t = e % W;       // [iu][iv]: the lanes of an item (iv) are contiguous
        const int iu = iu_t - shalf, iv = iv_t - shalf;
        CT out;
        out.x = out.y = (T)0;
        if (iu >= -(su / 2) && iu < su - su / 2 && iv >= -(sv / 2) && iv < sv - sv / 2) {
            const int cf_u = p.os_u * iu + (ou - p.os_u / 2 - 1) + p.n_cu / 2;
            const int cf_v = p.os_v * iv + (ov - p.os_v / 2 - 1) + p.n_cv / 2;
            if (cf_u >= 0 && cf_u < p.n_cu && cf_v >= 0 && cf_v < p.n_cv) {
                const double k = p.ck[((long long)cf * p.n_cu + cf_u) * p.n_cv + cf_v];
                const double2 g = p.pg[((long long)field * p.n_cu + cf_u) * p.n_cv + cf_v];
                out.x = (T)(k * g.x);
                out.y = (T)(k * g.y);
                sre += (double)out.x;
                sim += (double)out.y;
            }
        }
        tapsW[(long long)blockIdx.x * W * W + e] = out;
    }
    __shared__ double red[2][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sre += __shfl_xor_sync(0xffffffffu, sre, o);
        sim += __shfl_xor_sync(0xffffffffu, sim, o);
    }
    if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = sre, red[1][threadIdx.x >> 5] = sim;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < (int)((blockDim.x + 31) >> 5); ++i) a += red[0][i], b += red[1][i];
        tapnorm[blockIdx.x] = make_double2(a, b);
    }
}

template <typename T, int W, int PP> struct ApTrackCfg {
    static constexpr int IPW = 32 / W;
    static constexpr int ITER = W;
    // record: {uc | vc << 16, plane << 8 | flags} | tap block index | pad | PP (re, im) pairs of weighted data
    static constexpr int OFF_WD = 16;
    static constexpr int RAW = OFF_WD + PP * 2 * (int)sizeof(T);
    static constexpr int R16 = (RAW + 15) / 16;
    static constexpr int REC_BYTES = 16 * (R16 % 2 == 0 ? R16 + 1 : R16);   // 16 * odd: conflict-free 128-bit stores
    // tap-block ring: the W x W taps of a sample are ONE contiguous block of the tap table (2 KB for W = 16, fp32); the block
    // of the sample DEPTH steps ahead is fetched with cp.async.bulk (one instruction by one lane, completion on an mbarrier)
    // while the lanes multiply-accumulate the current one out of shared memory
    static constexpr int DEPTH = sizeof(T) == 4 ? 4 : 2;
    static constexpr int BLK_BYTES = W * W * 2 * (int)sizeof(T);
    static constexpr int RING_BYTES = DEPTH * IPW * BLK_BYTES;
    static constexpr int BAR_BYTES = DEPTH * IPW * 8;
    // per-lane sum_weight accumulators live in shared memory (updated once per staged sample): registers are what limits
    // this kernel's occupancy
    static constexpr int SW_BYTES = 32 * PP * 8;
    static constexpr int REC_TOTAL = 32 * REC_BYTES + SW_BYTES;
    static constexpr int WARP_BYTES = REC_TOTAL + RING_BYTES + BAR_BYTES;   // REC_TOTAL is a multiple of 16
};

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bulk_load(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    // arm the barrier with the byte count, then one bulk copy global -> shared that completes on it (UBLKCP)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
    } while (!done);
}

// blocks per SM the fp32 16 x 16 instantiation is compiled for: 3 = 168 registers, 12 warps per SM.  At 4 (128 registers, 16
// warps) round 1's kernel spilled 36 B inside the tap loop (5.31 ms against 2.87 ms on C3); round 2 moved the per-lane
// sum_weight accumulators to shared memory and re-reads the uv scales / recomputes the plane offsets instead of keeping
// them, which fits 128 registers without spills -- and is still slower (3.96 vs 3.08 ms through the wrapper): with fewer
// registers ptxas keeps fewer of the 16 tap loads of a sample in flight, and that costs more than the extra warps give.
#ifndef CNGI_AP_MINB_F32
#define CNGI_AP_MINB_F32 3
#endif
// BULK: taps come from the shared-memory ring filled by cp.async.bulk (see ApTrackCfg) instead of 16 L2 loads per sample
template <typename T, int W, int PP, bool BULK>
__global__ void __launch_bounds__(128, (sizeof(T) == 4 && W == 16) ? CNGI_AP_MINB_F32 : 1) aperture_track_kernel(ApParams p)
{
    using Cfg = ApTrackCfg<T, W, PP>;
    using CT = typename Cplx<T>::type;
    constexpr int IPW = Cfg::IPW, ITER = Cfg::ITER;
    const unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long task = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (task >= p.n_tasks) return;
    unsigned char *wbuf = smem + warp * (BULK ? Cfg::WARP_BYTES : Cfg::REC_TOTAL);
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(wbuf + Cfg::REC_TOTAL);
    const unsigned bar_s = ring_s + Cfg::RING_BYTES;
    if constexpr (BULK) {
        if (lane < Cfg::DEPTH * IPW) mbar_init(bar_s + lane * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }
    unsigned ring_parity = 0;   // BULK: bit `slot` = parity to wait for on this item's barrier of that slot
    const int S = p.smax, shalf = p.smax / 2;
    const int n_cf = p.n_cfb * p.n_cfc * p.n_cfp;

    const int cspan = (int)(task % p.n_cspan);
    long long rest = task / p.n_cspan;
    const int pgrp = (int)(rest % p.n_pgrp);
    rest /= p.n_pgrp;
    const int b = (int)(rest % p.n_baseline);
    const int seg = (int)(rest / p.n_baseline);
    const int t_lo = seg * p.seg_len, t_hi = min(p.n_time, t_lo + p.seg_len);
    const int G = p.G, spr = ITER >> p.log2G;
    const int c_base = cspan * IPW * G;
    const int p0 = pgrp * PP, npol = min(PP, p.n_pol - p0);
    const long long plane_cells = (long long)p.n_u * p.n_v;
    const int cf_b = (int)p.cf_b_map[b];

    // ---- phase-1 role: lane <-> staged sample ----
    const int k1 = lane % IPW, q1 = lane / IPW;
    const int g1 = q1 & (G - 1), row1 = q1 >> p.log2G;
    const int c1 = c_base + k1 * G + g1;
    const bool chan_ok = c1 < p.n_chan;
    const int a_chan1 = chan_ok ? ap_chan_of(p, c1) : 0;
    int cf1[PP];   // CF of each polarisation of the group (they usually coincide: cf_pol_map is often all zero)
#pragma unroll
    for (int ip = 0; ip < PP; ++ip)
        cf1[ip] = (chan_ok && ip < npol) ? (cf_b * p.n_cfc + (int)p.cf_c_map[c1]) * p.n_cfp + (int)p.cf_p_map[p0 + ip] : 0;
    double *sw_acc = reinterpret_cast<double *>(wbuf + 32 * Cfg::REC_BYTES) + lane * PP;   // shared memory, this lane's slots
#pragma unroll
    for (int ip = 0; ip < PP; ++ip) sw_acc[ip] = 0.0;
    long long carry_key = -1;

    // ---- phase-2 role: lane <-> (item, v residue mod W): lane r owns the grid row v == r (mod W) of the register window
    // (v is the contiguous grid axis, so a column leaving the window is reduced by W lanes into W consecutive cells),
    // accumulator j is the column u == j (mod W) ----
    const int k2 = lane / W, r2 = lane & (W - 1);
    int apol[PP];
#pragma unroll
    for (int ip = 0; ip < PP; ++ip) apol[ip] = (ip < npol) ? (p.pol_map ? (int)p.pol_map[p0 + ip] : p0 + ip) : 0;
    CT acc[W][PP];
#pragma unroll
    for (int j = 0; j < W; ++j)
#pragma unroll
        for (int ip = 0; ip < PP; ++ip) acc[j][ip].x = acc[j][ip].y = (T)0;
    int cur_plane = -1, lo_u = 0, lo_v = 0;

    auto flush_one = [&](int j, int u, int v) {
        const int cell = u * p.n_v + v;
#pragma unroll
        for (int ip = 0; ip < PP; ++ip) {
            if (ip < npol && (acc[j][ip].x != (T)0 || acc[j][ip].y != (T)0)) {
                red_add((CT *)p.grid + ((long long)cur_plane * p.n_ip + apol[ip]) * plane_cells + cell, acc[j][ip]);
                acc[j][ip].x = acc[j][ip].y = (T)0;   // clear only what was flushed (see standard_grid.cu)
            }
        }
    };
    auto my_row = [&]() { return lo_v + ((r2 - lo_v) & (W - 1)); };
    auto flush_row = [&]() {
        const int v = my_row();
#pragma unroll
        for (int j = 0; j < W; ++j) flush_one(j, lo_u + ((j - lo_u) & (W - 1)), v);
    };

    for (int t0 = t_lo; t0 < t_hi; t0 += spr) {
        // ---- phase 1 ----
        {
            unsigned char *rec = wbuf + lane * Cfg::REC_BYTES;
            int2 idx = make_int2(-1, 0);
            int blk[2] = {0, 0};
            long long key = -1;
            const int t = t0 + row1;
            if (chan_ok && t < t_hi) {
                const long long tb = (long long)t * p.n_baseline + b;
                const long long f = p.field[tb];
                const int field_indx = f > -1 ? ap_find_field(p, f) : -1;
                CellPos cp;
                bool ok = field_indx >= 0 && locate_centre(p.uvw[tb * 3], p.uvw[tb * 3 + 1], p.scale[c1], p.scale[p.n_chan + c1],
                                                           p.n_u, p.n_v, cp);   // (uv scales re-read per round: L1 hits, 4 registers saved)
                if (ok) ok = stamp_inside(cp.uc, cp.vc, p.max_support, p.n_u, p.n_v);
                if (ok) {
                    const long long s = (tb * p.n_chan + c1) * p.n_pol + p0;
                    T wd[PP * 2];
                    double wsel[PP];
                    bool any = false;
#pragma unroll
                    for (int ip = 0; ip < PP; ++ip) {
                        wd[2 * ip] = wd[2 * ip + 1] = (T)0;
                        wsel[ip] = 0.0;
                        if (ip < npol) {
                            const double w = (double)((const T *)p.weight)[s + ip];
                            double wre = w, wim = 0.0;
                            if (!p.do_psf) {
                                const CT d = ((const CT *)p.vis)[s + ip];
                                weighted_vis((double)d.x, (double)d.y, w, wre, wim);
                                if (p.flag && p.flag[s + ip]) wre = nan("");
                            }
                            if (!masked(wre, wim)) {
                                any = true;
                                wsel[ip] = w;
                                wd[2 * ip] = (T)wre, wd[2 * ip + 1] = (T)wim;
                            }
                        }
                    }
                    if (any) {
                        const int ou = oversample_offset(cp.uc, cp.u_pos, p.os_u) + p.os_u / 2 + 1;
                        const int ov = oversample_offset(cp.vc, cp.v_pos, p.os_v) + p.os_v / 2 + 1;
#pragma unroll
                        for (int ip = 0; ip < PP; ++ip) {
                            blk[ip] = (((field_indx * n_cf) + cf1[ip]) * p.n_off_u + ou) * p.n_off_v + ov;
                            const double2 tn = p.tapnorm[blk[ip]];
                            // psf: w * Re(norm); image: w * Re(norm^2)   (_aperture_grid.py:508-511)
                            sw_acc[ip] += wsel[ip] * (p.do_psf ? tn.x : (tn.x * tn.x - tn.y * tn.y));
                        }
#pragma unroll
                        for (int ip = 0; ip < PP; ++ip) {
                            CT v2;
                            v2.x = wd[2 * ip], v2.y = wd[2 * ip + 1];
                            reinterpret_cast<CT *>(rec + Cfg::OFF_WD)[ip] = v2;
                        }
                        idx = make_int2(cp.uc | (cp.vc << 16), a_chan1 << 8);
                        key = ((long long)a_chan1 * p.n_u + cp.uc) * p.n_v + cp.vc;
                    }
                }
            }
            long long prev = __shfl_up_sync(FULL, key, IPW);
            if (lane < IPW) prev = carry_key;
            carry_key = __shfl_sync(FULL, key, 32 - IPW + k1);
            if (key >= 0 && key == prev) idx.y |= 1;
            *reinterpret_cast<int4 *>(rec) = make_int4(idx.x, idx.y, blk[0], blk[PP - 1]);
        }
        __syncwarp();
        // ---- phase 2 ----
        auto prefetch = [&](int i) {   // the item's first lane fetches the tap block of step i into ring slot i mod DEPTH
            if (r2 == 0) {
                const int4 nx = *reinterpret_cast<const int4 *>(wbuf + (i * IPW + k2) * Cfg::REC_BYTES);
                if (nx.x != -1) {
                    const int slot = (i & (Cfg::DEPTH - 1)) * IPW + k2;
                    bulk_load(ring_s + slot * Cfg::BLK_BYTES, (const CT *)p.taps + (long long)nx.z * (W * W), Cfg::BLK_BYTES,
                              bar_s + slot * 8);
                }
            }
        };
        if constexpr (BULK) {
#pragma unroll
            for (int i = 0; i < Cfg::DEPTH; ++i) prefetch(i);
        }
#pragma unroll 1
        for (int i = 0; i < ITER; ++i) {
            const unsigned char *rec = wbuf + (i * IPW + k2) * Cfg::REC_BYTES;
            const int4 idx = *reinterpret_cast<const int4 *>(rec);
            if (idx.x != -1) {
            CT wd[PP];
#pragma unroll
            for (int ip = 0; ip < PP; ++ip) wd[ip] = reinterpret_cast<const CT *>(rec + Cfg::OFF_WD)[ip];
            const int uc = idx.x & 0xffff, vc = (int)((unsigned)idx.x >> 16);
            if (!(idx.y & 1)) {
                const int need_u = uc - shalf, need_v = vc - shalf;
                const int plane = idx.y >> 8;
                const bool new_plane = plane != cur_plane;
                int new_u = lo_u, new_v = lo_v;
                if (new_plane || need_u < lo_u) new_u = need_u;
                else if (need_u + S > lo_u + W) new_u = need_u + S - W;
                if (new_plane || need_v < lo_v) new_v = need_v;
                else if (need_v + S > lo_v + W) new_v = need_v + S - W;
                const int v = my_row();
                const bool row_leaves = new_plane ? (cur_plane >= 0) : (v < new_v || v >= new_v + W);
                if (row_leaves) {
                    flush_row();
                } else if (new_u != lo_u) {
#pragma unroll
                    for (int j = 0; j < W; ++j) {
                        const int u = lo_u + ((j - lo_u) & (W - 1));
                        if (u < new_u || u >= new_u + W) flush_one(j, u, v);
                    }
                }
                if (new_plane) cur_plane = plane;
                lo_u = new_u, lo_v = new_v;
            }
            // taps: slot s of this lane's row <-> stamp column qu = (s - bu) mod W, stamp row q = (r2 - bv) mod W
            const int bu = (uc - shalf) & (W - 1), bv = (vc - shalf) & (W - 1);
            const CT *tp = (const CT *)p.taps + (long long)idx.z * (W * W) + ((r2 - bv) & (W - 1));
            const CT *tp1 = (const CT *)p.taps + (long long)idx.w * (W * W) + ((r2 - bv) & (W - 1));
            const bool same_cf = idx.z == idx.w;
            CT d0[PP], d1[PP];   // acc += t.x * (dr, di) + t.y * (-di, dr)
#pragma unroll
            for (int ip = 0; ip < PP; ++ip) {
                d0[ip] = wd[ip];
                d1[ip].x = -wd[ip].y, d1[ip].y = wd[ip].x;
            }
            if constexpr (BULK) {
                const int slot = (i & (Cfg::DEPTH - 1)) * IPW + k2;
                mbar_wait(bar_s + slot * 8, (ring_parity >> (i & (Cfg::DEPTH - 1))) & 1u);
                ring_parity ^= 1u << (i & (Cfg::DEPTH - 1));
                const unsigned ts = ring_s + slot * Cfg::BLK_BYTES + ((r2 - bv) & (W - 1)) * (int)sizeof(CT);
#pragma unroll
                for (int sl = 0; sl < W; ++sl) {
                    const int row = ((sl - bu) & (W - 1)) * W;
                    CT tap;
                    if constexpr (sizeof(T) == 4)
                        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(tap.x), "=f"(tap.y) : "r"(ts + row * (int)sizeof(CT)));
                    else
                        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(tap.x), "=d"(tap.y) : "r"(ts + row * (int)sizeof(CT)));
#pragma unroll
                    for (int ip = 0; ip < PP; ++ip) {
                        if (ip > 0 && !same_cf) tap = tp1[row];   // a second convolution function for this pol: from L2
                        pair_fma(acc[sl][ip], d0[ip], tap.x);
                        pair_fma(acc[sl][ip], d1[ip], tap.y);
                    }
                }
            } else {
#pragma unroll
            for (int sl = 0; sl < W; ++sl) {
                const int row = ((sl - bu) & (W - 1)) * W;
                CT tap = tp[row];
#pragma unroll
                for (int ip = 0; ip < PP; ++ip) {
                    if (ip > 0 && !same_cf) tap = tp1[row];
                    pair_fma(acc[sl][ip], d0[ip], tap.x);
                    pair_fma(acc[sl][ip], d1[ip], tap.y);
                }
            }
            }
            }
            if constexpr (BULK) {
                __syncwarp();   // every lane has finished reading this step's ring slot: it may be refilled
                if (i + Cfg::DEPTH < ITER) prefetch(i + Cfg::DEPTH);
            }
        }
        __syncwarp();
    }
    if (cur_plane >= 0) flush_row();

    const int span = IPW * G;
#pragma unroll
    for (int ip = 0; ip < PP; ++ip) {
        double v = sw_acc[ip];
        for (int o = span; o < 32; o <<= 1) v += __shfl_xor_sync(FULL, v, o);
        warp_grouped_add(p.sum_weight, a_chan1 * p.n_ip + apol[ip], v, (lane < span) && chan_ok && (ip < npol));
    }
}

// A6 pass 1: bucket the weights
template <typename T> __global__ void __launch_bounds__(256) aperture_weight_bucket_kernel(ApParams p)
{
    const long long total = (long long)p.n_time * p.n_baseline * p.n_chan;
    const long long idx0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = idx0 < total;   // no early return: the warp aggregates its bucket updates together
    const long long idx = in_range ? idx0 : 0;
    const int c = (int)(idx % p.n_chan);
    const long long tb = idx / p.n_chan;
    const int b = (int)(tb % p.n_baseline);
    int field_indx = 0;
    CellPos cp;
    const bool ok = in_range && ap_locate(p, tb, c, field_indx, cp);
    const int cf_b = (int)p.cf_b_map[b], cf_c = (int)p.cf_c_map[c];
    const int a_chan = ap_chan_of(p, c);
    for (int ip = 0; ip < p.n_pol; ++ip) {   // uniform trip count
        double w = 0.0;
        bool use = ok;
        if (use) {
            w = (double)((const T *)p.weight)[idx * p.n_pol + ip];
            use = !(isnan(w) || w == 0.0);
        }
        const int cf_p = (int)p.cf_p_map[ip];
        const int a_pol = p.pol_map ? (int)p.pol_map[ip] : ip;
        const long long cf = (((long long)field_indx * p.n_cfb + cf_b) * p.n_cfc + cf_c) * p.n_cfp + cf_p;
        // neighbouring channels of a row share their bucket: one reduction per distinct bucket per warp
        warp_grouped_add(p.buckets, (int)((cf * p.n_ic + a_chan) * p.n_ip + a_pol), w, use);
    }
}

// A6 pass 2: one block per bucket; stamp bucket_weight * CF * PG at the grid centre (:276-289)
template <typename T> __global__ void __launch_bounds__(256) aperture_weight_stamp_kernel(ApParams p)
{
    using CT = typename Cplx<T>::type;
    const long long bucket = blockIdx.x;
    const double w = p.buckets[bucket];
    if (w == 0.0) return;
    long long r = bucket;
    const int a_pol = (int)(r % p.n_ip);
    r /= p.n_ip;
    const int a_chan = (int)(r % p.n_ic);
    r /= p.n_ic;
    const long long cf = r % ((long long)p.n_cfb * p.n_cfc * p.n_cfp);
    const int field_indx = (int)(r / ((long long)p.n_cfb * p.n_cfc * p.n_cfp));
    const int su = (int)p.support[cf * 2], sv = (int)p.support[cf * 2 + 1];
    const double *ck = p.ck + cf * p.n_cu * p.n_cv;
    const double2 *pg = p.pg + (long long)field_indx * p.n_cu * p.n_cv;
    CT *plane = (CT *)p.grid + ((long long)a_chan * p.n_ip + a_pol) * p.n_u * (long long)p.n_v;
    double nre = 0.0;
    for (int e = threadIdx.x; e < su * sv; e += blockDim.x) {
        const int iu = e / sv - su / 2, iv = e % sv - sv / 2;
        const int cf_u = p.os_u * iu + p.n_cu / 2, cf_v = p.os_v * iv + p.n_cv / 2;
        const double k = ck[(long long)cf_u * p.n_cv + cf_v];
        const double2 g = pg[(long long)cf_u * p.n_cv + cf_v];
        const double cr = k * g.x, ci = k * g.y;
        CT val;
        val.x = (T)(cr * w);
        val.y = (T)(ci * w);
        red_add(plane + (long long)(p.n_u / 2 + iu) * p.n_v + p.n_v / 2 + iv, val);
        nre += cr;
    }
    __shared__ double part[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nre += __shfl_xor_sync(0xffffffffu, nre, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = nre;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += part[i];
        atomicAdd(p.sum_weight + a_chan * p.n_ip + a_pol, w * s);   // sum over samples of w*Re(norm) == (sum w)*Re(norm)
    }
}

static int fill(ApParams &p, const cngi_aperture_grid_args *a, const char *who, bool need_vis)
{
    CNGI_REQUIRE(a != nullptr, "%s: null args", who);
    CNGI_REQUIRE(a->weight && a->uvw && a->freq_chan && a->field && a->field_id && a->cf_baseline_map && a->cf_chan_map &&
                     a->cf_pol_map && a->conv_kernel && a->weight_support && a->phase_gradient && a->grid && a->sum_weight,
                 "%s: null array pointer", who);
    CNGI_REQUIRE(!need_vis || a->vis, "%s: vis is null in image mode", who);
    CNGI_REQUIRE(a->precision == CNGI_F32 || a->precision == CNGI_F64, "%s: bad precision", who);
    CNGI_REQUIRE(a->chan_mode != CNGI_CHAN_GENERAL || a->chan_map, "%s: chan_map is null", who);
    CNGI_REQUIRE(a->n_u > 0 && a->n_v > 0 && a->n_u < (1 << 24) && a->n_v < (1 << 24), "%s: bad grid size", who);
    CNGI_REQUIRE(a->n_field > 0 && a->n_cfb > 0 && a->n_cfc > 0 && a->n_cfp > 0 && a->n_cu > 0 && a->n_cv > 0, "%s: bad CF shape", who);
    CNGI_REQUIRE(a->max_support >= 1, "%s: max_support must be >= 1", who);
    CNGI_REQUIRE(a->n_time * a->n_baseline < (1LL << 31), "%s: too many rows", who);
    // every tap index os*i + off + centre must stay inside the CF
    CNGI_REQUIRE((int64_t)a->oversampling[0] * (a->max_support / 2) + a->oversampling[0] / 2 + 1 <= a->n_cu / 2 + (a->n_cu % 2) &&
                     (int64_t)a->oversampling[1] * (a->max_support / 2) + a->oversampling[1] / 2 + 1 <= a->n_cv / 2 + (a->n_cv % 2),
                 "%s: CF of %lld x %lld is too small for support %d at oversampling %d x %d", who, (long long)a->n_cu,
                 (long long)a->n_cv, a->max_support, a->oversampling[0], a->oversampling[1]);
    p.n_time = (int)a->n_time, p.n_baseline = (int)a->n_baseline, p.n_chan = (int)a->n_chan, p.n_pol = (int)a->n_pol;
    p.n_ic = (int)a->n_imag_chan, p.n_ip = (int)a->n_imag_pol, p.n_u = (int)a->n_u, p.n_v = (int)a->n_v;
    p.vis = a->vis, p.weight = a->weight, p.flag = a->flag, p.uvw = a->uvw, p.freq = a->freq_chan;
    p.chan_map = a->chan_map, p.pol_map = a->pol_map, p.field = a->field, p.field_id = a->field_id;
    p.cf_b_map = a->cf_baseline_map, p.cf_c_map = a->cf_chan_map, p.cf_p_map = a->cf_pol_map, p.support = a->weight_support;
    p.ck = a->conv_kernel, p.pg = (const double2 *)a->phase_gradient, p.grid = a->grid, p.sum_weight = a->sum_weight;
    p.dl = a->delta_lm[0], p.dm = a->delta_lm[1];
    p.n_field = (int)a->n_field, p.n_cfb = (int)a->n_cfb, p.n_cfc = (int)a->n_cfc, p.n_cfp = (int)a->n_cfp;
    p.n_cu = (int)a->n_cu, p.n_cv = (int)a->n_cv, p.os_u = a->oversampling[0], p.os_v = a->oversampling[1];
    p.max_support = a->max_support, p.do_psf = a->do_psf, p.chan_mode = a->chan_mode;
    p.buckets = nullptr;
    return CNGI_OK;
}

template <typename T, int W, int PP> static int launch_aperture_track(ApParams p, cudaStream_t st)
{
    using CT = typename Cplx<T>::type;
    using Cfg = ApTrackCfg<T, W, PP>;
    p.smax = p.max_support;
    p.n_off_u = p.os_u + 3, p.n_off_v = p.os_v + 3;
    const int n_cf = p.n_cfb * p.n_cfc * p.n_cfp;
    const long long n_blocks = (long long)p.n_field * n_cf * p.n_off_u * p.n_off_v;
    CNGI_REQUIRE(n_blocks < (1LL << 31) && n_blocks * W * W * (long long)sizeof(CT) < (8LL << 30),
                 "aperture_grid: tap table of %lld blocks is too large", n_blocks);
    double *scale = nullptr;
    CT *taps = nullptr;
    double2 *tapnorm = nullptr;
    int rc = make_uv_scale_table(p.freq, p.n_chan, p.dl, p.dm, p.n_u, p.n_v, st, &scale);
    if (rc != CNGI_OK) return rc;
    tune_pool_once();
    cudaError_t e = cudaMallocAsync((void **)&taps, (size_t)n_blocks * W * W * sizeof(CT), st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&tapnorm, (size_t)n_blocks * sizeof(double2), st);
    if (e != cudaSuccess) {
        cudaFreeAsync(scale, st);
        if (taps) cudaFreeAsync(taps, st);
        set_error("aperture_grid: cudaMallocAsync of the tap table failed: %s", cudaGetErrorString(e));
        return CNGI_ERR_CUDA;
    }
    p.scale = scale, p.taps = taps, p.tapnorm = tapnorm;
    aperture_build_blocks_kernel<T, W><<<(unsigned)n_blocks, W * W < 256 ? W * W : 256, 0, st>>>(p, taps, tapnorm);
    // work decomposition (as the standard track kernel)
    int G = (p.chan_mode == CNGI_CHAN_CONTINUUM) ? 8 : 1;
    if (G > Cfg::ITER) G = Cfg::ITER;
    while (G > 1 && (Cfg::IPW * G / 2) >= p.n_chan) G >>= 1;
    int log2G = 0;
    while ((1 << (log2G + 1)) <= G) ++log2G;
    p.G = 1 << log2G, p.log2G = log2G;
    const int spr = Cfg::ITER / p.G;
    p.n_cspan = (int)ceil_div(p.n_chan, Cfg::IPW * p.G);
    p.n_pgrp = (int)ceil_div(p.n_pol, PP);
    const long long per_seg = (long long)p.n_baseline * p.n_cspan * p.n_pgrp;
    long long n_seg = ceil_div((long long)sm_count() * 16 * 12, per_seg);
    if (n_seg < 1) n_seg = 1;
    int seg_len = (int)ceil_div(p.n_time, n_seg);
    if (seg_len < 8 * spr) seg_len = 8 * spr;
    seg_len = (int)(ceil_div(seg_len, spr) * spr);
    p.seg_len = seg_len, p.n_seg = (int)ceil_div(p.n_time, seg_len);
    p.n_tasks = per_seg * p.n_seg;
    const int wpb = 4;
    const long long blocks = ceil_div(p.n_tasks, wpb);
    if (blocks >= (1LL << 31)) {
        cudaFreeAsync(taps, st), cudaFreeAsync(tapnorm, st), cudaFreeAsync(scale, st);
        set_error("aperture_grid: too many work items for one launch");
        return CNGI_ERR_INVALID;
    }
    {
        // Taps through the cp.async.bulk + mbarrier ring (CNGI_APERTURE_BULK=1) or straight from L2 (default).  Measured on
        // C3 (B200, fp32, 23.1 M samples): 3.21 ms with the ring, 3.08 ms without -- the 14.5 MB tap table is L2 resident and
        // its latency was already hidden; the ring's LDS + barrier waits cost more issue slots than the 16 LDGs they replace
        // (ncu: 54 % vs 46 % issue-active, profiles/r02_aperture_track*.txt).  The ring stays selectable for CFs that outgrow L2.
        static const bool bulk = getenv("CNGI_APERTURE_BULK") != nullptr;
        const int smem = wpb * (bulk ? Cfg::WARP_BYTES : Cfg::REC_TOTAL);
        auto kern = bulk ? aperture_track_kernel<T, W, PP, true> : aperture_track_kernel<T, W, PP, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) {
            kern<<<(unsigned)blocks, wpb * 32, smem, st>>>(p);
            e = cudaGetLastError();
        }
    }
    cudaFreeAsync(taps, st), cudaFreeAsync(tapnorm, st), cudaFreeAsync(scale, st);
    CNGI_CUDA_TRY(e);
    return CNGI_OK;
}

// fallback for supports above 15: offset-major taps, column-parallel lanes, reductions per tap
template <typename T> static int launch_aperture_rows(ApParams p, cudaStream_t st)
{
    using CT = typename Cplx<T>::type;
    const long long total = (long long)p.n_time * p.n_baseline * p.n_chan;
    p.smax = p.max_support;
    p.n_off_u = p.os_u + 3, p.n_off_v = p.os_v + 3;
    const int n_cf = p.n_cfb * p.n_cfc * p.n_cfp;
    const long long n_elems = (long long)p.n_field * n_cf * p.n_off_u * p.n_off_v * p.smax * p.smax;
    CNGI_REQUIRE(p.smax <= 32, "aperture_grid: max support above 32 is not supported (got %d)", p.smax);
    CNGI_REQUIRE(n_elems * (long long)sizeof(CT) < (8LL << 30), "aperture_grid: offset-major tap table would need %lld MB",
                 n_elems * (long long)sizeof(CT) >> 20);
    double *scale = nullptr;
    int rc = make_uv_scale_table(p.freq, p.n_chan, p.dl, p.dm, p.n_u, p.n_v, st, &scale);
    if (rc != CNGI_OK) return rc;
    p.scale = scale;
    CT *taps = nullptr;
    tune_pool_once();
    cudaError_t e = cudaMallocAsync((void **)&taps, (size_t)n_elems * sizeof(CT), st);
    if (e != cudaSuccess) {
        cudaFreeAsync(scale, st);
        set_error("aperture_grid: cudaMallocAsync of the tap table failed: %s", cudaGetErrorString(e));
        return CNGI_ERR_CUDA;
    }
    aperture_build_taps_kernel<T><<<(unsigned)std::min<long long>(ceil_div(n_elems, 256), (long long)sm_count() * 16), 256, 0, st>>>(p, taps, n_elems);
    p.taps = taps;
    long long blocks = ceil_div(ceil_div(total, 32) * 32, 256);
    const long long cap = (long long)sm_count() * 8 * 4;
    if (blocks > cap) blocks = cap;
    if (p.smax <= 8)
        aperture_grid_kernel<T, 8><<<(unsigned)blocks, 256, 0, st>>>(p);
    else if (p.smax <= 16)
        aperture_grid_kernel<T, 16><<<(unsigned)blocks, 256, 0, st>>>(p);
    else
        aperture_grid_kernel<T, 32><<<(unsigned)blocks, 256, 0, st>>>(p);
    e = cudaGetLastError();
    cudaFreeAsync(taps, st);
    cudaFreeAsync(scale, st);
    CNGI_CUDA_TRY(e);
    return CNGI_OK;
}

template <typename T> static int launch_aperture(ApParams p, cudaStream_t st, int algorithm)
{
    CNGI_REQUIRE(p.n_u < 65536 && p.n_v < 65536 && p.os_u < 65000 && p.os_v < 65000, "aperture_grid: grid side / oversampling too large");
    CNGI_REQUIRE(p.n_cfb * p.n_cfc < (1 << 20) && p.n_field < (1 << 11), "aperture_grid: too many convolution functions / fields");
    CNGI_REQUIRE((long long)p.n_u * p.n_v < (1LL << 31), "aperture_grid: n_u*n_v overflows int32");
    const bool track = algorithm != CNGI_ALGO_NAIVE && p.max_support <= 15 && p.n_ic < (1 << 23);
    if (!track) return launch_aperture_rows<T>(p, st);
    if (p.max_support <= 7)
        return p.n_pol == 1 ? launch_aperture_track<T, 8, 1>(p, st) : launch_aperture_track<T, 8, 2>(p, st);
    return p.n_pol == 1 ? launch_aperture_track<T, 16, 1>(p, st) : launch_aperture_track<T, 16, 2>(p, st);
}

}  // namespace cngi

extern "C" int cngi_b200_aperture_grid(const cngi_aperture_grid_args *a, void *stream)
{
    using namespace cngi;
    ApParams p{};
    int rc = fill(p, a, "aperture_grid", a && !a->do_psf);
    if (rc != CNGI_OK) return rc;
    if (p.do_psf) p.flag = nullptr;
    const long long total = (long long)p.n_time * p.n_baseline * p.n_chan;
    if (total == 0 || p.n_pol == 0) return CNGI_OK;
    const int algo = getenv("CNGI_APERTURE_NAIVE") ? CNGI_ALGO_NAIVE : CNGI_ALGO_AUTO;   // development knob
    return a->precision == CNGI_F32 ? launch_aperture<float>(p, (cudaStream_t)stream, algo)
                                    : launch_aperture<double>(p, (cudaStream_t)stream, algo);
}

extern "C" int cngi_b200_aperture_weight_grid(const cngi_aperture_grid_args *a, void *stream)
{
    using namespace cngi;
    ApParams p{};
    int rc = fill(p, a, "aperture_weight_grid", false);
    if (rc != CNGI_OK) return rc;
    const long long total = (long long)p.n_time * p.n_baseline * p.n_chan;
    if (total == 0 || p.n_pol == 0) return CNGI_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const long long n_buckets = (long long)p.n_field * p.n_cfb * p.n_cfc * p.n_cfp * p.n_ic * p.n_ip;
    CNGI_REQUIRE(n_buckets < (1LL << 31), "aperture_weight_grid: too many (field, cf, plane) buckets");
    double *scale = nullptr;
    rc = make_uv_scale_table(p.freq, p.n_chan, p.dl, p.dm, p.n_u, p.n_v, st, &scale);
    if (rc != CNGI_OK) return rc;
    p.scale = scale;
    tune_pool_once();
    CNGI_CUDA_TRY(cudaMallocAsync((void **)&p.buckets, n_buckets * sizeof(double), st));
    CNGI_CUDA_TRY(cudaMemsetAsync(p.buckets, 0, n_buckets * sizeof(double), st));
    const long long blocks = ceil_div(total, 256);
    CNGI_REQUIRE(blocks < (1LL << 31), "aperture_weight_grid: too many samples");
    if (a->precision == CNGI_F32) {
        aperture_weight_bucket_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(p);
        aperture_weight_stamp_kernel<float><<<(unsigned)n_buckets, 256, 0, st>>>(p);
    } else {
        aperture_weight_bucket_kernel<double><<<(unsigned)blocks, 256, 0, st>>>(p);
        aperture_weight_stamp_kernel<double><<<(unsigned)n_buckets, 256, 0, st>>>(p);
    }
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(p.buckets, st);
    cudaFreeAsync(scale, st);
    CNGI_CUDA_TRY(e);
    return CNGI_OK;
}
