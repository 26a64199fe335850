// standard_degrid.cu -- A7 of SURVEY.md section 8: degridding predict, the adjoint of the standard gridder.
// The reference has no implementation (ngcasa/imaging/predict_modelvis_image.py:20-40 is a docstring stub and
// _standard_grid.py:418-430 prints "still needs to be implemented"), so the specification is: same cell /
// offset / bounds arithmetic as _standard_grid_jit (_standard_grid.py:299-324), gather instead of scatter.
//
// S lanes cooperate on one (time, baseline, chan) sample: lane i reads the grid row u = uc - S/2 + i (S cells,
// contiguous in v, shared by neighbouring samples through L1/L2), weights it with the v taps, multiplies by its u
// tap, and the S partial sums are combined with shuffles.  All polarisations reuse the taps.
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
    const double *scale;   // [2, n_chan] uv_scale table
};

// generic support: `LW` lanes per sample (power of two >= support, <= 32)
template <typename T, int LW> __global__ void __launch_bounds__(256) std_degrid_kernel(DgParams p)
{
    using CT = typename Cplx<T>::type;
    constexpr int SPW = 32 / LW;   // samples per warp
    const unsigned FULL = 0xffffffffu;
    const long long total = (long long)p.n_time * p.n_baseline * p.n_chan;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LW, li = lane % LW;
    long long idx = warp * SPW + sub;
    const bool in_range = idx < total;
    if (!in_range) idx = total - 1;   // keep the warp converged for the shuffles
    const int c = (int)(idx % p.n_chan);
    const long long tb = idx / p.n_chan;
    const int half = p.support / 2;
    CellPos cp;
    bool ok = locate_centre(p.uvw[tb * 3], p.uvw[tb * 3 + 1], p.scale[c], p.scale[p.n_chan + c], p.n_u, p.n_v, cp);
    if (ok) ok = stamp_inside(cp.uc, cp.vc, half, p.n_u, p.n_v);
    int uoff = 0, voff = 0;
    if (ok) {
        uoff = oversample_offset(cp.uc, cp.u_pos, p.oversampling);
        voff = oversample_offset(cp.vc, cp.v_pos, p.oversampling);
    }
    const int a_chan = p.chan_mode == CNGI_CHAN_CUBE ? c : (p.chan_mode == CNGI_CHAN_CONTINUUM ? 0 : (int)p.chan_map[c]);
    const bool lane_on = ok && li < p.support;
    const double cu = lane_on ? p.cgk[abs(p.oversampling * (li - half) + uoff)] : 0.0;
    double inv_norm = 1.0;
    if (p.normalize) {   // norm = sum over the stamp of cu*cv = (sum cu) * (sum cv)
        double su = cu, sv = 0.0;
#pragma unroll
        for (int o = LW / 2; o > 0; o >>= 1) su += __shfl_xor_sync(FULL, su, o);
        if (ok)
            for (int q = 0; q < p.support; ++q) sv += p.cgk[abs(p.oversampling * (q - half) + voff)];
        inv_norm = ok ? 1.0 / (su * sv) : 1.0;
    }
    for (int ip = 0; ip < p.n_pol; ++ip) {
        double are = 0.0, aim = 0.0;
        if (lane_on) {
            const int a_pol = p.pol_map ? (int)p.pol_map[ip] : ip;
            const CT *row = (const CT *)p.grid + (((long long)a_chan * p.n_ip + a_pol) * p.n_u + cp.uc - half + li) * p.n_v +
                            cp.vc - half;
            for (int q = 0; q < p.support; ++q) {
                const double cv = p.cgk[abs(p.oversampling * (q - half) + voff)];
                const CT g = row[q];
                are += cv * (double)g.x;
                aim += cv * (double)g.y;
            }
            are *= cu;
            aim *= cu;
        }
#pragma unroll
        for (int o = LW / 2; o > 0; o >>= 1) {
            are += __shfl_xor_sync(FULL, are, o);
            aim += __shfl_xor_sync(FULL, aim, o);
        }
        if (in_range && li == 0) {
            CT out;
            out.x = (T)(are * inv_norm);
            out.y = (T)(aim * inv_norm);
            ((CT *)p.vis)[idx * p.n_pol + ip] = out;   // skipped samples reach here with 0
        }
    }
}

template <typename T> static int launch_degrid(DgParams p, cudaStream_t st)
{
    const long long total = (long long)p.n_time * p.n_baseline * p.n_chan;
    int lw = 1;
    while (lw < p.support) lw <<= 1;
    const int spw = 32 / lw;
    const long long warps = ceil_div(total, spw);
    const long long blocks = ceil_div(warps * 32, 256);
    CNGI_REQUIRE(blocks < (1LL << 31), "standard_degrid: too many samples");
    double *scale = nullptr;
    int rc = make_uv_scale_table(p.freq, p.n_chan, p.dl, p.dm, p.n_u, p.n_v, st, &scale);
    if (rc != CNGI_OK) return rc;
    p.scale = scale;
    switch (lw) {
        case 1: std_degrid_kernel<T, 1><<<(unsigned)blocks, 256, 0, st>>>(p); break;
        case 2: std_degrid_kernel<T, 2><<<(unsigned)blocks, 256, 0, st>>>(p); break;
        case 4: std_degrid_kernel<T, 4><<<(unsigned)blocks, 256, 0, st>>>(p); break;
        case 8: std_degrid_kernel<T, 8><<<(unsigned)blocks, 256, 0, st>>>(p); break;
        case 16: std_degrid_kernel<T, 16><<<(unsigned)blocks, 256, 0, st>>>(p); break;
        default: std_degrid_kernel<T, 32><<<(unsigned)blocks, 256, 0, st>>>(p); break;
    }
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(scale, st);
    CNGI_CUDA_TRY(e);
    return CNGI_OK;
}

}  // namespace cngi

extern "C" int cngi_b200_standard_degrid(const cngi_std_degrid_args *a, void *stream)
{
    using namespace cngi;
    CNGI_REQUIRE(a != nullptr, "standard_degrid: null args");
    CNGI_REQUIRE(a->model_grid && a->uvw && a->freq_chan && a->cgk_1D && a->vis, "standard_degrid: null array pointer");
    CNGI_REQUIRE(a->precision == CNGI_F32 || a->precision == CNGI_F64, "standard_degrid: bad precision");
    CNGI_REQUIRE(a->support >= 1 && a->support <= 32 && a->oversampling >= 0, "standard_degrid: support must be in [1, 32]");
    CNGI_REQUIRE(a->chan_mode != CNGI_CHAN_GENERAL || a->chan_map, "standard_degrid: chan_map is null");
    CNGI_REQUIRE(a->n_u > 0 && a->n_v > 0 && a->n_u < (1 << 24) && a->n_v < (1 << 24), "standard_degrid: bad grid size");
    CNGI_REQUIRE(a->n_time * a->n_baseline < (1LL << 31), "standard_degrid: too many rows");
    if (a->n_time == 0 || a->n_baseline == 0 || a->n_chan == 0 || a->n_pol == 0) return CNGI_OK;
    DgParams p{};
    p.n_time = (int)a->n_time, p.n_baseline = (int)a->n_baseline, p.n_chan = (int)a->n_chan, p.n_pol = (int)a->n_pol;
    p.n_ic = (int)a->n_imag_chan, p.n_ip = (int)a->n_imag_pol, p.n_u = (int)a->n_u, p.n_v = (int)a->n_v;
    p.grid = a->model_grid, p.uvw = a->uvw, p.freq = a->freq_chan, p.chan_map = a->chan_map, p.pol_map = a->pol_map;
    p.cgk = a->cgk_1D, p.vis = a->vis, p.dl = a->delta_lm[0], p.dm = a->delta_lm[1];
    p.support = a->support, p.oversampling = a->oversampling, p.chan_mode = a->chan_mode, p.normalize = a->normalize;
    return a->precision == CNGI_F32 ? launch_degrid<float>(p, (cudaStream_t)stream) : launch_degrid<double>(p, (cudaStream_t)stream);
}
