// fft_bluestein.cuh -- shared-memory FFTs of length n1 * P (P prime in (512, 1021]); see fft_bluestein.cu
#pragma once
#include "common.cuh"

namespace cngi {

struct BluAxis {             // device tables of one (length, direction)
    int n = 0, n1 = 0, P = 0;
    float2 *tw_n = nullptr, *w_n1 = nullptr, *chirp = nullptr, *chirp_out = nullptr, *bf = nullptr, *w_m = nullptr;
};

// n = n1 * P with P prime, 512 < P <= 1021, n1 <= 32 (measured: cuFFT is faster below -- 1228 = 4 * 307: 0.05 vs 0.15 ms)
bool blu_supported(int64_t n, int *n1_out = nullptr, int *p_out = nullptr);
int blu_axis_create(BluAxis *ax, int64_t n, int sign);   // sign -1: forward (exp(-i..)), +1: unnormalised inverse
void blu_axis_destroy(BluAxis *ax);
// transforms n_lines x n_planes lines; element j of line i of plane q is at q * plane + i * line + j * elem (in elements)
int blu_lines(const BluAxis &ax, const float2 *src, float2 *dst, long long src_line, long long src_elem, long long src_plane,
              long long dst_line, long long dst_elem, long long dst_plane, int n_lines, int n_planes, cudaStream_t st,
              int line0 = 0, int line_mod = 0,    // lines (line0 + i) mod line_mod, i < n_lines (a cyclic window of lines)
              const float *src_real = nullptr);   // real input lines (src ignored): same strides, read as (x, 0)

}  // namespace cngi
