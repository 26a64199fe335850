// standard_grid_window16_f64.cu -- the register-window gridder (standard_grid_window.cuh) with 16-wide windows: supports
// 9 / 11 / 13 / 15 of the standard gridder (the reference is support-generic, _standard_grid.py:344-360), complex128 / float64.
// 16 lanes per item, two items per warp, 16 x 16 cells per item in registers; the tap table stores every row twice and only
// the sub-vector rotations as copies (see WinCfg).  fp64 keeps one polarisation per item (64 accumulator registers);
// the other polarisations are further work items.
#include "standard_grid_window.cuh"

namespace cngi {

template <bool CPLX, int S> static int launch16_f64_s(StdParams p, const cngi_std_grid_args *a, cudaStream_t st)
{
    return launch_window_t<double, CPLX, S, 1, false>(p, a, st);
}

int launch_window16_f64(StdParams p, const cngi_std_grid_args *a, cudaStream_t st)
{
#define CNGI_W16_CASE(SS) \
    case SS: return a->complex_grid ? launch16_f64_s<true, SS>(p, a, st) : launch16_f64_s<false, SS>(p, a, st);
    switch (a->support) {
        CNGI_W16_CASE(9)
        CNGI_W16_CASE(11)
        CNGI_W16_CASE(13)
        CNGI_W16_CASE(15)
    }
#undef CNGI_W16_CASE
    set_error("standard_grid: the 16-wide window kernel handles supports 9, 11, 13 and 15 (got %d)", a->support);
    return CNGI_ERR_UNSUPPORTED;
}

}  // namespace cngi
