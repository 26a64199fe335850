// standard_grid_window.cu -- instantiations of the register-window gridder (standard_grid_window.cuh) for the 8-wide windows:
// supports 3 / 5 / 7, the fused image + psf pass and the fused imaging-weight pass.
#include "standard_grid_window.cuh"

namespace cngi {

template <typename T, bool CPLX, int S> static int launch_window_pp(StdParams p, const cngi_std_grid_args *a, cudaStream_t st)
{
    static const int nz_knob = env_knob("CNGI_WIN_NZ", -1);
    // single-channel items on a grid far larger than L2 are bound by the reductions: skip never-touched cells there
    const double grid_bytes = (double)p.n_ic * p.n_ip * p.n_u * p.n_v * sizeof(T) * (CPLX ? 2 : 1);
    bool nz = p.chan_mode != CNGI_CHAN_CONTINUUM && grid_bytes > 4e9;
    if (nz_knob >= 0) nz = nz_knob != 0;
    if (p.n_pol == 1) return nz ? launch_window_t<T, CPLX, S, 1, true>(p, a, st) : launch_window_t<T, CPLX, S, 1, false>(p, a, st);
    return nz ? launch_window_t<T, CPLX, S, 2, true>(p, a, st) : launch_window_t<T, CPLX, S, 2, false>(p, a, st);
}

template <typename T, bool CPLX> static int launch_window_s(StdParams p, const cngi_std_grid_args *a, cudaStream_t st)
{
#ifdef CNGI_WIN_MINIMAL   // SASS experiments: instantiate one kernel only
    return launch_window_t<float, true, 7, 2, false>(p, a, st);
#else
    switch (a->support) {
        case 3: return launch_window_pp<T, CPLX, 3>(p, a, st);
        case 5: return launch_window_pp<T, CPLX, 5>(p, a, st);
        default: return launch_window_pp<T, CPLX, 7>(p, a, st);
    }
#endif
}

bool window_kernel_supported(const cngi_std_grid_args *a, int table_len)
{
    if (a->support < 3 || a->support > 15 || a->support % 2 == 0) return false;
    if (a->oversampling < 1 || table_len > 8192) return false;
    const int w = a->support < 4 ? 4 : a->support < 8 ? 8 : 16;
    if (a->n_u < w || a->n_v < w || a->n_u > 32767 || a->n_v > 32767) return false;   // packed 16-bit cell keys
    // the tap table (W rotations of W taps; 16-wide windows: 16 / sizeof(vector) rotations of 2 x 16) + one double per
    // oversampling offset: kept under ~56 KB so that 3-4 blocks fit an SM
    const int n_off = a->oversampling + 3;
    const int tsz = a->precision == CNGI_F32 ? 4 : 8;
    const long long row_table = w <= 8 ? (long long)w * w * tsz : (long long)(16 / tsz) * 32 * tsz;
    return (long long)n_off * (row_table + 8) <= 56 * 1024;
}

// fused image + psf pass (support 7, the one make_image / make_psf use: make_image.py:106-107)
int launch_window_dual(StdParams p, const cngi_std_grid_args *a, cudaStream_t st)
{
#ifdef CNGI_WIN_MINIMAL
    return CNGI_ERR_UNSUPPORTED;
#else
    if (a->precision == CNGI_F32)
        return p.n_pol == 1 ? launch_window_t<float, true, 7, 1, false, true>(p, a, st)
                            : launch_window_t<float, true, 7, 2, false, true>(p, a, st);
    return p.n_pol == 1 ? launch_window_t<double, true, 7, 1, false, true>(p, a, st)
                        : launch_window_t<double, true, 7, 2, false, true>(p, a, st);
#endif
}

// imaging weights formed inside the gridder (support 7, complex image grid: what make_image runs after make_imaging_weight)
int launch_window_iw(StdParams p, const cngi_std_grid_args *a, cudaStream_t st)
{
#ifdef CNGI_WIN_MINIMAL
    return CNGI_ERR_UNSUPPORTED;
#else
    if (a->precision == CNGI_F32)
        return p.n_pol == 1 ? launch_window_t<float, true, 7, 1, false, false, true>(p, a, st)
                            : launch_window_t<float, true, 7, 2, false, false, true>(p, a, st);
    return p.n_pol == 1 ? launch_window_t<double, true, 7, 1, false, false, true>(p, a, st)
                        : launch_window_t<double, true, 7, 2, false, false, true>(p, a, st);
#endif
}

int launch_window(StdParams p, const cngi_std_grid_args *a, cudaStream_t st)
{
    if (a->support > 8) return a->precision == CNGI_F32 ? launch_window16_f32(p, a, st) : launch_window16_f64(p, a, st);
    if (a->precision == CNGI_F32)
        return a->complex_grid ? launch_window_s<float, true>(p, a, st) : launch_window_s<float, false>(p, a, st);
    return a->complex_grid ? launch_window_s<double, true>(p, a, st) : launch_window_s<double, false>(p, a, st);
}

}  // namespace cngi
