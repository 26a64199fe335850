// standard_degrid.cu -- A7 of SURVEY.md section 8: degridding predict, the adjoint of the standard gridder.
// The reference has no implementation (ngcasa/imaging/predict_modelvis_image.py:20-40 is a docstring stub and
// _standard_grid.py:418-430 prints "still needs to be implemented"), so the specification is: same cell /
// offset / bounds arithmetic as _standard_grid_jit (_standard_grid.py:299-324), gather instead of scatter.
//
// Kernel shape (gather, L1-wavefront bound):
//   * a warp takes 32 consecutive (time, baseline, chan) samples per round.  Step 1: lane L does the fp64 bit-exact
//     cell / offset arithmetic of sample L (no redundancy).  Step 2: LW lanes cooperate on one sample, lane j owning
//     grid COLUMN v = vc - S/2 + j -- the fastest axis -- so the LW lanes of a sample read one contiguous run of a
//     grid row (one 128-byte line) per load instead of LW scattered rows; each lane accumulates
//     sum_i cu[i] * G[uc - S/2 + i][v_j] over the S rows, multiplies by its own v tap, and the partial sums are
//     combined with log2(LW) shuffles.  All polarisations reuse the taps.
//   * taps come from a shared-memory table laid out [offset][tap] (one 128-bit row load gives a sample's S taps).
#include "common.cuh"

namespace cngi {

struct DgParams {
    int n_time, n_baseline, n_chan, n_pol;
    int n_ic, n_ip, n_u, n_v;
    const void *grid;
    const double *uvw;
    const double *freq;
    const int64_t *chan_map, *pol_map;
    const double *cgk;
    void *vis;
    double dl, dm;
    int support, oversampling, chan_mode, normalize;
    int table_len;
    const double *scale;   // [2, n_chan] uv_scale table
};

// LW = lanes per sample (power of two >= support), SP = taps per table row (support rounded up to LW)
template <typename T, int LW, int NP> __global__ void __launch_bounds__(256) std_degrid_kernel(DgParams p)
{
    using CT = typename Cplx<T>::type;
    constexpr int SPW = 32 / LW;   // samples processed concurrently by a warp
    const unsigned FULL = 0xffffffffu;
    const int n_pol = NP ? NP : p.n_pol;
    const int half = p.support / 2;

    // tap rows: taps[(off + os/2 + 1) * LW + q] = cgk[|os*(q - half) + off|] for q < support, else 0;
    // tsum[off + os/2 + 1] = sum of that row
    extern __shared__ __align__(16) unsigned char smem[];
    T *taps = reinterpret_cast<T *>(smem);
    const int n_off = p.oversampling + 3;
    T *tsum = taps + n_off * LW;
    for (int e = threadIdx.x; e < n_off * LW; e += blockDim.x) {
        const int q = e % LW, off = e / LW - p.oversampling / 2 - 1;
        const int k = min(abs(p.oversampling * (q - half) + off), p.table_len - 1);
        taps[e] = q < p.support ? (T)p.cgk[k] : (T)0;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < n_off; o += blockDim.x) {
        T sum = (T)0;
        for (int q = 0; q < p.support; ++q) sum += taps[o * LW + q];
        tsum[o] = sum;
    }
    __syncthreads();

    const long long total = (long long)p.n_time * p.n_baseline * p.n_chan;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LW, lj = lane % LW;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int o0 = p.oversampling / 2 + 1;
    const long long plane_cells = (long long)p.n_u * p.n_v;

    for (long long base = warp0 * 32; base < total; base += n_warps * 32) {
        // ---- step 1: lane L locates sample base + L ------------------------------------------------------------
        const long long mine = base + lane;
        int pack0 = -1, pack1 = 0, my_plane = 0;   // {uc | vc << 16}, {uoff | voff << 16} (table row indices), image plane
        if (mine < total) {
            const int c = (int)(mine % p.n_chan);
            const long long tb = mine / p.n_chan;
            CellPos cp;
            bool ok = locate_centre(p.uvw[tb * 3], p.uvw[tb * 3 + 1], p.scale[c], p.scale[p.n_chan + c], p.n_u, p.n_v, cp);
            if (ok) ok = stamp_inside(cp.uc, cp.vc, half, p.n_u, p.n_v);
            if (ok) {
                pack0 = cp.uc | (cp.vc << 16);
                pack1 = (oversample_offset(cp.uc, cp.u_pos, p.oversampling) + o0) |
                        ((oversample_offset(cp.vc, cp.v_pos, p.oversampling) + o0) << 16);
                my_plane = p.chan_mode == CNGI_CHAN_CUBE ? c : (p.chan_mode == CNGI_CHAN_CONTINUUM ? 0 : (int)p.chan_map[c]);
            }
        }
        // ---- step 2: LW lanes per sample, lane lj <-> grid column vc - half + lj ---------------------------------------
#pragma unroll 1
        for (int it = 0; it < LW; ++it) {
            const int src = it * SPW + sub;                       // which of the 32 located samples this lane group takes
            const int q0 = __shfl_sync(FULL, pack0, src);
            const int q1 = __shfl_sync(FULL, pack1, src);
            const int plane = __shfl_sync(FULL, my_plane, src);
            const long long idx = base + src;
            const bool ok = q0 != -1;
            const bool on = ok && lj < p.support;
            const int uc = q0 & 0xffff, vc = (int)((unsigned)q0 >> 16);
            const int ou = q1 & 0xffff, ov = (int)((unsigned)q1 >> 16);
            const T *ru = taps + ou * LW;
            const T cv = on ? taps[ov * LW + lj] : (T)0;
            const T inv = (ok && p.normalize) ? (T)1 / (tsum[ou] * tsum[ov]) : (T)1;   // 1 / (sum of the S*S taps)
            const long long cell0 = on ? (long long)(uc - half) * p.n_v + vc - half + lj : 0;
#pragma unroll
            for (int ip = 0; ip < n_pol; ++ip) {   // uniform trip count (compile-time when NP != 0)
                const int a_pol = (!NP && p.pol_map) ? (int)p.pol_map[ip] : ip;
                T sre = (T)0, sim = (T)0;
                if (on) {
                    const CT *g = (const CT *)p.grid + ((long long)plane * p.n_ip + a_pol) * plane_cells + cell0;
                    for (int i = 0; i < p.support; ++i) {
                        const CT x = g[(long long)i * p.n_v];
                        const T cu = ru[i];
                        sre = fma(cu, x.x, sre);
                        sim = fma(cu, x.y, sim);
                    }
                }
                sre *= cv;
                sim *= cv;
#pragma unroll
                for (int o = LW / 2; o > 0; o >>= 1) {
                    sre += __shfl_xor_sync(FULL, sre, o);
                    sim += __shfl_xor_sync(FULL, sim, o);
                }
                if (lj == 0 && idx < total) {
                    CT o2;
                    o2.x = ok ? sre * inv : (T)0;   // samples the gridder would skip yield exactly 0
                    o2.y = ok ? sim * inv : (T)0;
                    ((CT *)p.vis)[idx * n_pol + ip] = o2;
                }
            }
        }
    }
}

template <typename T, int LW> static int launch_degrid_np(const DgParams &p, int np, unsigned blocks, size_t smem, cudaStream_t st)
{
    if (np == 1) std_degrid_kernel<T, LW, 1><<<blocks, 256, smem, st>>>(p);
    else if (np == 2) std_degrid_kernel<T, LW, 2><<<blocks, 256, smem, st>>>(p);
    else std_degrid_kernel<T, LW, 0><<<blocks, 256, smem, st>>>(p);
    CNGI_CUDA_TRY(cudaGetLastError());
    return CNGI_OK;
}

template <typename T> static int launch_degrid(DgParams p, bool identity_pol, cudaStream_t st)
{
    const long long total = (long long)p.n_time * p.n_baseline * p.n_chan;
    int lw = 1;
    while (lw < p.support) lw <<= 1;
    long long blocks = ceil_div(ceil_div(total, 32) * 32, 256);
    const long long cap = (long long)sm_count() * 8 * 4;   // grid-stride loop: a few waves of resident blocks
    if (blocks > cap) blocks = cap;
    const size_t smem = (size_t)(p.oversampling + 3) * (lw + 1) * sizeof(T);
    CNGI_REQUIRE(smem <= 48 * 1024, "standard_degrid: tap table too large for shared memory (%zu bytes)", smem);
    const int np = (identity_pol && (p.n_pol == 1 || p.n_pol == 2)) ? p.n_pol : 0;
    double *scale = nullptr;
    int rc = make_uv_scale_table(p.freq, p.n_chan, p.dl, p.dm, p.n_u, p.n_v, st, &scale);
    if (rc != CNGI_OK) return rc;
    p.scale = scale;
    switch (lw) {
        case 1: rc = launch_degrid_np<T, 1>(p, np, (unsigned)blocks, smem, st); break;
        case 2: rc = launch_degrid_np<T, 2>(p, np, (unsigned)blocks, smem, st); break;
        case 4: rc = launch_degrid_np<T, 4>(p, np, (unsigned)blocks, smem, st); break;
        case 8: rc = launch_degrid_np<T, 8>(p, np, (unsigned)blocks, smem, st); break;
        case 16: rc = launch_degrid_np<T, 16>(p, np, (unsigned)blocks, smem, st); break;
        default: rc = launch_degrid_np<T, 32>(p, np, (unsigned)blocks, smem, st); break;
    }
    cudaFreeAsync(scale, st);
    return rc;
}

// standard_degrid_window.cu
bool degrid_window_supported(const cngi_std_degrid_args *a);
int launch_degrid_window(const cngi_std_degrid_args *a, cudaStream_t st);

}  // namespace cngi

extern "C" int cngi_b200_standard_degrid(const cngi_std_degrid_args *a, void *stream)
{
    using namespace cngi;
    CNGI_REQUIRE(a != nullptr, "standard_degrid: null args");
    CNGI_REQUIRE(a->model_grid && a->uvw && a->freq_chan && a->cgk_1D && a->vis, "standard_degrid: null array pointer");
    CNGI_REQUIRE(a->precision == CNGI_F32 || a->precision == CNGI_F64, "standard_degrid: bad precision");
    CNGI_REQUIRE(a->support >= 1 && a->support <= 32 && a->oversampling >= 0, "standard_degrid: support must be in [1, 32]");
    CNGI_REQUIRE(a->chan_mode != CNGI_CHAN_GENERAL || a->chan_map, "standard_degrid: chan_map is null");
    CNGI_REQUIRE(a->n_u > 0 && a->n_v > 0 && a->n_u < 65536 && a->n_v < 65536, "standard_degrid: grid side must be below 65536");
    CNGI_REQUIRE(a->oversampling < 65000, "standard_degrid: oversampling too large");
    CNGI_REQUIRE(a->n_time * a->n_baseline < (1LL << 31), "standard_degrid: too many rows");
    if (a->n_time == 0 || a->n_baseline == 0 || a->n_chan == 0 || a->n_pol == 0) return CNGI_OK;
    CNGI_REQUIRE(a->algorithm >= 0 && a->algorithm <= 2, "standard_degrid: bad algorithm %d", a->algorithm);
    if (a->algorithm == 2 && !degrid_window_supported(a)) {
        set_error("standard_degrid: window kernel needs support in {3,5,7}, 1 or 2 pols with the identity pol_map");
        return CNGI_ERR_UNSUPPORTED;
    }
    if (a->algorithm != 1 && degrid_window_supported(a)) return launch_degrid_window(a, (cudaStream_t)stream);
    DgParams p{};
    p.n_time = (int)a->n_time, p.n_baseline = (int)a->n_baseline, p.n_chan = (int)a->n_chan, p.n_pol = (int)a->n_pol;
    p.n_ic = (int)a->n_imag_chan, p.n_ip = (int)a->n_imag_pol, p.n_u = (int)a->n_u, p.n_v = (int)a->n_v;
    p.grid = a->model_grid, p.uvw = a->uvw, p.freq = a->freq_chan, p.chan_map = a->chan_map, p.pol_map = a->pol_map;
    p.cgk = a->cgk_1D, p.vis = a->vis, p.dl = a->delta_lm[0], p.dm = a->delta_lm[1];
    p.support = a->support, p.oversampling = a->oversampling, p.chan_mode = a->chan_mode, p.normalize = a->normalize;
    p.table_len = a->oversampling * (a->support / 2 + 1);
    if (p.table_len < 1) p.table_len = 1;
    const bool identity_pol = a->pol_map == nullptr;
    return a->precision == CNGI_F32 ? launch_degrid<float>(p, identity_pol, (cudaStream_t)stream)
                                    : launch_degrid<double>(p, identity_pol, (cudaStream_t)stream);
}
