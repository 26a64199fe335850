// standard_grid_window16_f32.cu -- the register-window gridder (standard_grid_window.cuh) with 16-wide windows: supports
// 9 / 11 / 13 / 15 of the standard gridder (the reference is support-generic, _standard_grid.py:344-360), complex64 / float32.
// 16 lanes per item, two items per warp, 16 x 16 cells per item in registers; the tap table stores every row twice and only
// the sub-vector rotations as copies (see WinCfg).
#include "standard_grid_window.cuh"

namespace cngi {

template <bool CPLX, int S> static int launch16_f32_s(StdParams p, const cngi_std_grid_args *a, cudaStream_t st)
{
    if (p.n_pol == 1) return launch_window_t<float, CPLX, S, 1, false>(p, a, st);
    return launch_window_t<float, CPLX, S, 2, false>(p, a, st);
}

int launch_window16_f32(StdParams p, const cngi_std_grid_args *a, cudaStream_t st)
{
#define CNGI_W16_CASE(SS) \
    case SS: return a->complex_grid ? launch16_f32_s<true, SS>(p, a, st) : launch16_f32_s<false, SS>(p, a, st);
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
