// fft_bluestein.cu -- complex64 FFTs of length n = n1 * P (P a prime in (512, 1021], n1 <= 32) done per line in SHARED MEMORY.
//
// Why: the reference pads the uv-grid to int(1.2 * image_size) (_check_imaging_parms.py:36) -- for the power-of-two images
// people make that is 1228 = 4 * 307, 4915 = 5 * 983, 9830 = 10 * 983 (BASELINE config 5), 19660 = 20 * 983.  cuFFT has no
// native radix for such primes and falls back to a full-length Bluestein: measured on a B200, 6.9 ms per 9830^2 complex64
// plane against 1.8 ms for the friendly 10240^2 -- which made the FFT 93 % of a config-5 cube step (profiles/r02_*).  Here
// the line is split n = n1 * P (Cooley-Tukey): n1-point DFTs + twiddles on the line held in shared memory, then n1
// Bluestein transforms of length P whose chirp convolutions are 2048-point FFTs that never leave shared memory
// (radix 8 x 8 x 8 x 4 decimation in frequency forward, the mirrored decimation in time backward -- the two digit
// reversals cancel, so nothing is reordered; the last forward stage, the multiplication with the filter spectrum and the
// first backward stage are one register-resident step).  One block per line, one global read and one global write of the
// line.  A 2-D transform is two passes of the same kernel: rows (contiguous lines), then columns (strided lines, in place).
//
// cuFFT remains the transform for every other size and for complex128 (cngi_b200_fft_plan_create decides); results of the
// two paths agree to fp32 rounding (tests/test_gpu_fft.py).
#include "fft_bluestein.cuh"
#include <cmath>
#include <complex>
#include <vector>

namespace cngi {

namespace {

constexpr int BM = 2048;                      // convolution length (>= 2 P - 1)
constexpr int BT = 256;                       // threads per block
constexpr int YLEN = BM + (BM >> 5) * 4;      // complex work buffer, skewed by 4 elements per 32

// 64-bit shared-memory accesses are served per half warp: the skew makes the 16 elements a half warp touches in the
// stride-4 stage (4 blocks of 32 x 4 offsets) fall into 16 different 8-byte bank pairs
__device__ __forceinline__ int skew(int e) { return e + ((e >> 5) << 2); }

#ifndef CNGI_BLU_PACKED
#define CNGI_BLU_PACKED 1
#endif
#if CNGI_BLU_PACKED
// complex add / subtract as ONE packed instruction (add.rn.f32x2 / sub.rn.f32x2, sm_100+): the butterflies' additions were
// 33 % of the kernel's instructions as scalar FADDs (ncu, profiles/r02_bluestein_9830.txt)
__device__ __forceinline__ float2 operator+(float2 a, float2 b)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 operator-(float2 a, float2 b)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tsub.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
#else
__device__ __forceinline__ float2 operator+(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 operator-(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
#if CNGI_BLU_PACKED
// complex multiply (-accumulate) in TWO packed instructions: a * b = (ax, ay) * bx + (-ay, ax) * by.  ptxas folds the swapped,
// half-negated copy of `a` into an operand modifier of the second one (SASS: FMUL2 t, a, bx ; FFMA2 d, -a.LO_HI.NP, by, t) --
// no data movement, against 2 FMUL + 2 FFMA for the scalar form.
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    float2 r;
    asm("{.reg .b64 ra, rs, bx, by, t, rd;\n\t.reg .f32 nay;\n\t"
        "neg.f32 nay, %3;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rs, {nay, %2};\n\t"
        "mov.b64 bx, {%4, %4};\n\tmov.b64 by, {%5, %5};\n\t"
        "mul.rn.f32x2 t, ra, bx;\n\tfma.rn.f32x2 rd, rs, by, t;\n\tmov.b64 {%0, %1}, rd;}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
// acc + a * b
__device__ __forceinline__ float2 cmac(float2 acc, float2 a, float2 b)
{
    float2 r;
    asm("{.reg .b64 ra, rs, bx, by, t, rd;\n\t.reg .f32 nay;\n\t"
        "neg.f32 nay, %3;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rs, {nay, %2};\n\t"
        "mov.b64 bx, {%4, %4};\n\tmov.b64 by, {%5, %5};\n\tmov.b64 t, {%6, %7};\n\t"
        "fma.rn.f32x2 t, ra, bx, t;\n\tfma.rn.f32x2 rd, rs, by, t;\n\tmov.b64 {%0, %1}, rd;}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(acc.x), "f"(acc.y));
    return r;
}
#else
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cmac(float2 acc, float2 a, float2 b)
{
    return make_float2(fmaf(a.x, b.x, fmaf(-a.y, b.y, acc.x)), fmaf(a.x, b.y, fmaf(a.y, b.x, acc.y)));
}
#endif
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiplication by SIGN * i
template <int SIGN> __device__ __forceinline__ float2 rot90(float2 v) { return SIGN < 0 ? make_float2(v.y, -v.x) : make_float2(-v.y, v.x); }

// X_k = sum_j a_j W^{jk}, W = exp(SIGN * 2 pi i / 4), in place, natural order
template <int SIGN> __device__ __forceinline__ void dft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3)
{
    const float2 s02 = a0 + a2, d02 = a0 - a2, s13 = a1 + a3, d13 = rot90<SIGN>(a1 - a3);
    a0 = s02 + s13, a2 = s02 - s13, a1 = d02 + d13, a3 = d02 - d13;
}

// (x + i y) * exp(SIGN * 2 pi i / 8) and * exp(SIGN * 2 pi i * 3 / 8)
template <int SIGN> __device__ __forceinline__ float2 mul_w8_1(float2 v)
{
    const float r = 0.70710678118654752440f;
#if CNGI_BLU_PACKED
    return cmul(v, make_float2(r, SIGN * r));     // two packed instructions
#else
    return make_float2((v.x - SIGN * v.y) * r, (v.y + SIGN * v.x) * r);
#endif
}
template <int SIGN> __device__ __forceinline__ float2 mul_w8_3(float2 v)
{
    const float r = 0.70710678118654752440f;
#if CNGI_BLU_PACKED
    return cmul(v, make_float2(-r, SIGN * r));
#else
    return make_float2((-v.x - SIGN * v.y) * r, (-v.y + SIGN * v.x) * r);
#endif
}

// 8-point DFT in place, natural order in and out
template <int SIGN> __device__ __forceinline__ void dft8(float2 (&a)[8])
{
    float2 b0 = a[0] + a[4], b1 = a[1] + a[5], b2 = a[2] + a[6], b3 = a[3] + a[7];
    float2 c0 = a[0] - a[4], c1 = mul_w8_1<SIGN>(a[1] - a[5]), c2 = rot90<SIGN>(a[2] - a[6]), c3 = mul_w8_3<SIGN>(a[3] - a[7]);
    dft4<SIGN>(b0, b1, b2, b3);
    dft4<SIGN>(c0, c1, c2, c3);
    a[0] = b0, a[1] = c0, a[2] = b1, a[3] = c1, a[4] = b2, a[5] = c2, a[6] = b3, a[7] = c3;
}

// the same with a[4..7] == 0 (the zero-padded half of the chirp convolution's input)
template <int SIGN> __device__ __forceinline__ void dft8_low_half(float2 (&a)[8])
{
    float2 b0 = a[0], b1 = a[1], b2 = a[2], b3 = a[3];
    float2 c0 = a[0], c1 = mul_w8_1<SIGN>(a[1]), c2 = rot90<SIGN>(a[2]), c3 = mul_w8_3<SIGN>(a[3]);
    dft4<SIGN>(b0, b1, b2, b3);
    dft4<SIGN>(c0, c1, c2, c3);
    a[0] = b0, a[1] = c0, a[2] = b1, a[3] = c1, a[4] = b2, a[5] = c2, a[6] = b3, a[7] = c3;
}

// the three table values a stage's twiddles are derived from; they depend on the thread and the stage only, so the
// kernel loads them ONCE per line and keeps them in registers across the n1 Bluestein transforms
struct TwBase {
    float2 w1, w2, w4;
};
__device__ __forceinline__ TwBase tw_base(const float2 *__restrict__ wm, int j1)
{
    TwBase b;
    b.w1 = __ldg(wm + j1), b.w2 = __ldg(wm + 2 * j1), b.w4 = __ldg(wm + 4 * j1);
    return b;
}
template <bool CONJ> __device__ __forceinline__ void twiddles_from(const TwBase &b, float2 (&w)[8])
{
    float2 w1 = b.w1, w2 = b.w2, w4 = b.w4;
    if (CONJ) w1 = cconj(w1), w2 = cconj(w2), w4 = cconj(w4);
    w[1] = w1, w[2] = w2, w[4] = w4;
    w[3] = cmul(w1, w2), w[5] = cmul(w4, w1), w[6] = cmul(w4, w2), w[7] = cmul(w4, w[3]);
}

// N-point DFT of a register array with the table w[j] = W_NW^j (W_N = w[TS]): even sizes split once more into two
// half-size DFTs + N / 2 twiddle products (10 -> 2 x 5: 55 complex multiply-adds instead of 100), odd sizes directly
template <int N, int TS, int NW>
__device__ __forceinline__ void small_dft(const float2 (&in)[N], float2 (&out)[N], const float2 (&w)[NW])
{
    if constexpr (N == 1) {
        out[0] = in[0];
    } else if constexpr (N == 2) {
        out[0] = in[0] + in[1], out[1] = in[0] - in[1];
    } else if constexpr (N % 2 == 0) {
        float2 e[N / 2], o[N / 2], E[N / 2], O[N / 2];
#pragma unroll
        for (int m = 0; m < N / 2; ++m) e[m] = in[2 * m], o[m] = in[2 * m + 1];
        small_dft<N / 2, 2 * TS, NW>(e, E, w);
        small_dft<N / 2, 2 * TS, NW>(o, O, w);
#pragma unroll
        for (int k = 0; k < N / 2; ++k) {
            const float2 tw = k ? cmul(O[k], w[(k * TS) % NW]) : O[k];
            out[k] = E[k] + tw, out[k + N / 2] = E[k] - tw;
        }
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            float2 acc = in[0];
#pragma unroll
            for (int l = 1; l < N; ++l) {
                const float2 c = w[(l * k * TS) % NW];
                if ((l * k) % N == 0) {
                    acc = acc + in[l];
                } else {
                    acc = cmac(acc, in[l], c);
                }
            }
            out[k] = acc;
        }
    }
}

struct BluParams {
    const float2 *src;
    const float *src_real;     // when set, the lines are REAL (psf grids): read as (x, 0), same element strides
    float2 *dst;
    long long src_line, src_elem, src_plane, dst_line, dst_elem, dst_plane;   // element strides
    int n_lines, n1, P;
    int line0, line_mod;       // line handled by block i is (line0 + i) mod line_mod: a cyclic window of lines
    const float2 *tw_n;        // [n]   exp(s 2 pi i j / n)
    const float2 *w_n1;        // [n1]  exp(s 2 pi i j / n1)
    const float2 *chirp;       // [P]   exp(s i pi j^2 / P)
    const float2 *chirp_out;   // [P]   chirp / M
    const float2 *bf;          // [M]   FFT_M of the conjugate chirp, in the digit-reversed order the forward network leaves
    const float2 *w_m;         // [M]   exp(-2 pi i j / M)
};

// One radix-8 stage of the 2048-point network on the skewed re / im planes.  DIF (forward): butterfly, then twiddle;
// DIT (backward): twiddle, then butterfly.  `first` = element index of leg 0, `sub` = distance between legs, j1 = index of
// W^{offset} in the W_M table.
template <bool DIT> __device__ __forceinline__ void radix8_stage(float2 *y, const TwBase &tb, int first, int sub)
{
    float2 a[8], w[8];
#pragma unroll
    for (int l = 0; l < 8; ++l) a[l] = y[skew(first + sub * l)];
    twiddles_from<DIT>(tb, w);
    if (DIT) {
#pragma unroll
        for (int q = 1; q < 8; ++q) a[q] = cmul(a[q], w[q]);
        dft8<+1>(a);
    } else {
        dft8<-1>(a);
#pragma unroll
        for (int q = 1; q < 8; ++q) a[q] = cmul(a[q], w[q]);
    }
#pragma unroll
    for (int l = 0; l < 8; ++l) y[skew(first + sub * l)] = a[l];
}

// N1 > 0: compile-time n1 (the n1-point DFTs keep a column in registers); N1 == 0: any n1 <= 32 (re-reads shared memory)
template <int N1> __global__ void __launch_bounds__(BT, 2) bluestein_lines_kernel(BluParams p)
{
    extern __shared__ __align__(16) unsigned char blu_smem[];
    const int n1 = N1 > 0 ? N1 : p.n1, P = p.P, n = n1 * P, t = threadIdx.x;
    float2 *X = reinterpret_cast<float2 *>(blu_smem);              // the line: x[n1' * P + n2]
    float2 *y = X + ((n + 1) & ~1);                                 // convolution buffer, 16-byte aligned
    int line = p.line0 + (int)blockIdx.x;
    if (line >= p.line_mod) line -= p.line_mod;
    const float2 *__restrict__ src = p.src + (long long)blockIdx.y * p.src_plane + (long long)line * p.src_line;
    float2 *dst = p.dst +   // (may alias src: the column pass runs in place; every read of the line precedes its writes)
        (long long)blockIdx.y * p.dst_plane + (long long)line * p.dst_line;

    if (p.src_real) {
        const float *__restrict__ sr = p.src_real + (long long)blockIdx.y * p.src_plane + (long long)line * p.src_line;
#pragma unroll 8
        for (int j = t; j < n; j += BT) X[j] = make_float2(__ldg(sr + (long long)j * p.src_elem), 0.f);
    } else {
#pragma unroll 8
        for (int j = t; j < n; j += BT) X[j] = __ldg(src + (long long)j * p.src_elem);
    }
    __syncthreads();

    // ---- n1-point DFTs down the columns of the (n1, P) view + twiddles: A[k1][n2] = W_n^{n2 k1} sum_l x[l P + n2] W_n1^{l k1}
    for (int n2 = t; n2 < P; n2 += BT) {
        if constexpr (N1 > 0) {
            float2 x[N1], w[N1], o[N1];
#pragma unroll
            for (int l = 0; l < N1; ++l) x[l] = X[l * P + n2], w[l] = __ldg(p.w_n1 + l);
            small_dft<N1, 1, N1>(x, o, w);
#pragma unroll
            for (int k1 = 0; k1 < N1; ++k1) X[k1 * P + n2] = k1 ? cmul(o[k1], __ldg(p.tw_n + n2 * k1)) : o[k1];
        } else {
            float2 out[32];
            for (int k1 = 0; k1 < n1; ++k1) {
                float2 acc = X[n2];
                for (int l = 1; l < n1; ++l) acc = acc + cmul(X[l * P + n2], __ldg(p.w_n1 + (l * k1) % n1));
                out[k1] = k1 ? cmul(acc, __ldg(p.tw_n + n2 * k1)) : acc;
            }
            for (int k1 = 0; k1 < n1; ++k1) X[k1 * P + n2] = out[k1];
        }
    }
    __syncthreads();

    // ---- n1 Bluestein transforms of length P: B[k1][k2] = c[k2] * IFFT_M(FFT_M(A[k1] * c, zero padded) * bf)[k2]
    const TwBase tw2048 = tw_base(p.w_m, t), tw256 = tw_base(p.w_m, 8 * (t & 31)), tw32 = tw_base(p.w_m, 64 * (t & 3));
    float2 chirp[4];   // this thread's four chirp values (elements t + 256 l): the same for every k1; the output chirp is chirp / M
#pragma unroll
    for (int l = 0; l < 4; ++l) chirp[l] = (t + 256 * l < P) ? __ldg(p.chirp + t + 256 * l) : make_float2(0.f, 0.f);
    for (int k1 = 0; k1 < n1; ++k1) {
        const float2 *seg = X + k1 * P;
        {   // forward stage 1 (span 2048) fused with the load: legs 4..7 are the zero padding (P <= 1024)
            float2 a[8], w[8];
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                const int j = t + 256 * l;
                a[l] = j < P ? cmul(seg[j], chirp[l]) : make_float2(0.f, 0.f);
            }
            dft8_low_half<-1>(a);
            twiddles_from<false>(tw2048, w);
#pragma unroll
            for (int q = 1; q < 8; ++q) a[q] = cmul(a[q], w[q]);
#pragma unroll
            for (int q = 0; q < 8; ++q) y[skew(t + 256 * q)] = a[q];
        }
        __syncthreads();
        radix8_stage<false>(y, tw256, (t >> 5) * 256 + (t & 31), 32);     // span 256
        __syncthreads();
        radix8_stage<false>(y, tw32, (t >> 2) * 32 + (t & 3), 4);         // span 32
        __syncthreads();
        // forward span 4, times the filter spectrum, backward span 4: four consecutive elements, registers only
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int g = t + 256 * h, e = skew(4 * g);   // 4 g .. 4 g + 3 never straddle a skew boundary
            const float4 v01 = *reinterpret_cast<float4 *>(y + e), v23 = *reinterpret_cast<float4 *>(y + e + 2);
            float2 a0 = make_float2(v01.x, v01.y), a1 = make_float2(v01.z, v01.w), a2 = make_float2(v23.x, v23.y), a3 = make_float2(v23.z, v23.w);
            dft4<-1>(a0, a1, a2, a3);
            const float4 f01 = __ldg(reinterpret_cast<const float4 *>(p.bf + 4 * g));
            const float4 f23 = __ldg(reinterpret_cast<const float4 *>(p.bf + 4 * g + 2));
            a0 = cmul(a0, make_float2(f01.x, f01.y)), a1 = cmul(a1, make_float2(f01.z, f01.w));
            a2 = cmul(a2, make_float2(f23.x, f23.y)), a3 = cmul(a3, make_float2(f23.z, f23.w));
            dft4<+1>(a0, a1, a2, a3);
            *reinterpret_cast<float4 *>(y + e) = make_float4(a0.x, a0.y, a1.x, a1.y);
            *reinterpret_cast<float4 *>(y + e + 2) = make_float4(a2.x, a2.y, a3.x, a3.y);
        }
        __syncthreads();
        radix8_stage<true>(y, tw32, (t >> 2) * 32 + (t & 3), 4);          // span 32
        __syncthreads();
        radix8_stage<true>(y, tw256, (t >> 5) * 256 + (t & 31), 32);      // span 256
        __syncthreads();
        {   // backward span 2048 fused with the output chirp; only k2 < P (<= 1024: legs 0..3) is kept, into the dead segment
            float2 a[8], w[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] = y[skew(t + 256 * q)];
            twiddles_from<true>(tw2048, w);
#pragma unroll
            for (int q = 1; q < 8; ++q) a[q] = cmul(a[q], w[q]);
            dft8<+1>(a);
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                const int k2 = t + 256 * l;
                if (k2 < P) X[k1 * P + k2] = cmul(a[l], make_float2(chirp[l].x * (1.0f / BM), chirp[l].y * (1.0f / BM)));
            }
        }
        __syncthreads();
    }

    // ---- X[k1 * P + k2] is bin k1 + n1 * k2
    for (int j = t; j < n; j += BT) dst[(long long)j * p.dst_elem] = X[(j % n1) * P + j / n1];
}

int upload(const std::vector<std::complex<double>> &h, float2 **dev)
{
    std::vector<float2> f(h.size());
    for (size_t i = 0; i < h.size(); ++i) f[i] = make_float2((float)h[i].real(), (float)h[i].imag());
    CNGI_CUDA_TRY(cudaMalloc((void **)dev, f.size() * sizeof(float2)));
    CNGI_CUDA_TRY(cudaMemcpy(*dev, f.data(), f.size() * sizeof(float2), cudaMemcpyHostToDevice));
    return CNGI_OK;
}

void fft_pow2(std::vector<std::complex<double>> &a)   // in-place radix-2, forward, for the one-off filter spectrum
{
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    const double pi = std::acos(-1.0);
    for (size_t len = 2; len <= n; len <<= 1)
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const std::complex<double> w = std::polar(1.0, -2.0 * pi * (double)k / (double)len);
                const std::complex<double> u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v, a[i + k + len / 2] = u - v;
            }
}

}  // namespace

bool blu_supported(int64_t n, int *n1_out, int *p_out)
{
    if (n < 128 || n > 32 * 1021) return false;
    for (int n1 = 1; n1 <= 32; ++n1) {
        if (n % n1) continue;
        const int64_t P = n / n1;
        if (P <= 512 || P > 1021) continue;   // the 2048-point convolution is the right size for 512 < P <= 1021 only
        bool prime = true;
        for (int64_t d = 2; d * d <= P; ++d)
            if (P % d == 0) { prime = false; break; }
        if (prime) {
            if (n1_out) *n1_out = n1;
            if (p_out) *p_out = (int)P;
            return true;
        }
    }
    return false;
}

int blu_axis_create(BluAxis *ax, int64_t n, int sign)
{
    *ax = BluAxis{};
    if (!blu_supported(n, &ax->n1, &ax->P)) return CNGI_ERR_UNSUPPORTED;
    ax->n = (int)n;
    const int n1 = ax->n1, P = ax->P;
    const double pi = std::acos(-1.0), s = sign < 0 ? -1.0 : 1.0;
    std::vector<std::complex<double>> tw(n), w1(n1), c(P), cout(P), b(BM), wm(BM);
    for (int64_t j = 0; j < n; ++j) tw[j] = std::polar(1.0, s * 2.0 * pi * (double)j / (double)n);
    for (int j = 0; j < n1; ++j) w1[j] = std::polar(1.0, s * 2.0 * pi * (double)j / (double)n1);
    for (int64_t j = 0; j < P; ++j) {
        c[j] = std::polar(1.0, s * pi * (double)((j * j) % (2 * P)) / (double)P);   // j^2 reduced mod 2P exactly
        cout[j] = c[j] / (double)BM;
    }
    for (int j = 0; j < BM; ++j) b[j] = 0.0, wm[j] = std::polar(1.0, -2.0 * pi * (double)j / (double)BM);
    for (int j = 0; j < P; ++j) {
        b[j] = std::conj(c[j]);
        if (j) b[BM - j] = std::conj(c[j]);
    }
    fft_pow2(b);
    std::vector<std::complex<double>> bperm(BM);   // position p of the forward network holds frequency k1 + 8 (k2 + 8 (k3 + 8 k4))
    for (int p = 0; p < BM; ++p) {
        int q = p;
        const int k4 = q % 4; q /= 4;
        const int k3 = q % 8; q /= 8;
        const int k2 = q % 8; q /= 8;
        bperm[p] = b[q + 8 * (k2 + 8 * (k3 + 8 * k4))];
    }
    int rc;
    if ((rc = upload(tw, &ax->tw_n)) || (rc = upload(w1, &ax->w_n1)) || (rc = upload(c, &ax->chirp)) ||
        (rc = upload(cout, &ax->chirp_out)) || (rc = upload(bperm, &ax->bf)) || (rc = upload(wm, &ax->w_m))) {
        blu_axis_destroy(ax);
        return rc;
    }
    return CNGI_OK;
}

void blu_axis_destroy(BluAxis *ax)
{
    for (float2 **q : {&ax->tw_n, &ax->w_n1, &ax->chirp, &ax->chirp_out, &ax->bf, &ax->w_m}) {
        if (*q) cudaFree(*q);
        *q = nullptr;
    }
}

int blu_lines(const BluAxis &ax, const float2 *src, float2 *dst, long long src_line, long long src_elem, long long src_plane,
              long long dst_line, long long dst_elem, long long dst_plane, int n_lines, int n_planes, cudaStream_t st,
              int line0, int line_mod, const float *src_real)
{
    if (n_lines <= 0 || n_planes <= 0) return CNGI_OK;
    BluParams p{};
    p.src = src, p.dst = dst, p.src_real = src_real;
    p.src_line = src_line, p.src_elem = src_elem, p.src_plane = src_plane;
    p.dst_line = dst_line, p.dst_elem = dst_elem, p.dst_plane = dst_plane;
    p.n_lines = n_lines, p.n1 = ax.n1, p.P = ax.P;
    p.line0 = line0, p.line_mod = line_mod > 0 ? line_mod : n_lines;
    p.tw_n = ax.tw_n, p.w_n1 = ax.w_n1, p.chirp = ax.chirp, p.chirp_out = ax.chirp_out, p.bf = ax.bf, p.w_m = ax.w_m;
    const size_t smem = (size_t)((ax.n + 1) & ~1) * sizeof(float2) + YLEN * sizeof(float2);
    CNGI_REQUIRE(smem <= 227 * 1024, "fft: line of %d complex64 does not fit shared memory", ax.n);
    CNGI_REQUIRE(n_planes < 65536, "fft: too many planes per batch");
    const dim3 grid((unsigned)n_lines, (unsigned)n_planes);
#define CNGI_BLU_LAUNCH(K)                                                                                        \
    do {                                                                                                          \
        CNGI_CUDA_TRY(cudaFuncSetAttribute(bluestein_lines_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        bluestein_lines_kernel<K><<<grid, BT, smem, st>>>(p);                                                     \
    } while (0)
    switch (ax.n1) {
        case 2: CNGI_BLU_LAUNCH(2); break;
        case 4: CNGI_BLU_LAUNCH(4); break;
        case 5: CNGI_BLU_LAUNCH(5); break;
        case 10: CNGI_BLU_LAUNCH(10); break;
        case 20: CNGI_BLU_LAUNCH(20); break;
        default: CNGI_BLU_LAUNCH(0); break;
    }
#undef CNGI_BLU_LAUNCH
    CNGI_CUDA_TRY(cudaGetLastError());
    return CNGI_OK;
}

}  // namespace cngi
