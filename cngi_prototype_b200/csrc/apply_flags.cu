// apply_flags.cu -- N4: cngi.vis.apply_flags on the device (reference cngi/vis/apply_flags.py:53, the same masking
// synthesis_imaging_cube.py:180 does in place with vis_data[flag] = nan).
//
//   flagged[dv] = dv.where(FLAG == 0).astype(dv.dtype)      for every data variable with FLAG's dims
//
// xarray's where() fills with dtypes.get_fill_value(dtype): NaN for floats, NaN + NaN j for complex (xarray
// core/dtypes.py maybe_promote), so a flagged real becomes the quiet NaN 0x7ff8000000000000 (0x7fc00000 in float32) and
// a flagged complex (NaN, NaN).  The output is bit-identical to numpy.where(flag == 0, data, fill).astype(dtype).
//
// HBM-bound byte work.  In place (out == data) only the flag bytes are read -- 16 per thread with one 128-bit load --
// and only flagged elements are stored, so the algorithmic traffic is 1 B per element + elem_bytes per flagged element.
// Out of place the kernel is a copy with select: (2 * elem_bytes + 1) B per element.
#include "common.cuh"

namespace cngi {
namespace {

template <int KIND> struct Elem;
template <> struct Elem<CNGI_ELEM_F32> {
    using type = unsigned int;
    static __device__ __forceinline__ type nan() { return 0x7fc00000u; }
};
template <> struct Elem<CNGI_ELEM_F64> {
    using type = unsigned long long;
    static __device__ __forceinline__ type nan() { return 0x7ff8000000000000ull; }
};
template <> struct Elem<CNGI_ELEM_C64> {
    using type = uint2;
    static __device__ __forceinline__ type nan() { return make_uint2(0x7fc00000u, 0x7fc00000u); }
};
template <> struct Elem<CNGI_ELEM_C128> {
    using type = ulonglong2;
    static __device__ __forceinline__ type nan() { return make_ulonglong2(0x7ff8000000000000ull, 0x7ff8000000000000ull); }
};

__device__ __forceinline__ void count_flagged(int mine, unsigned long long *n_flagged)
{
    if (n_flagged == nullptr) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_flagged, (unsigned long long)mine);
}

// In place: flag bytes [head, head + 16 * n_vec) are 16-byte aligned; the first `head` and the last elements are
// handled one by one by the first block.  (8 blocks per SM through __launch_bounds__(256, 8) was measured: 32 registers
// with small spills, 0.115 ms instead of 0.109 ms on 115.6 M complex64 -- occupancy is not what limits it.)
template <int KIND>
__global__ void __launch_bounds__(256) apply_flags_inplace_kernel(typename Elem<KIND>::type *__restrict__ data,
                                                                  const unsigned char *__restrict__ flag, long long n,
                                                                  long long head, long long n_vec,
                                                                  unsigned long long *n_flagged)
{
    using E = Elem<KIND>;
    int mine = 0;
    const uint4 *fv = reinterpret_cast<const uint4 *>(flag + head);
    typename E::type *body = data + head;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_vec; i0 += 4 * stride) {
        uint4 f4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)   // four independent 16-byte flag loads in flight
            f4[q] = (i0 + q * stride < n_vec) ? __ldg(fv + i0 + q * stride) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 f = f4[q];
            if ((f.x | f.y | f.z | f.w) == 0u) continue;   // the common case: nothing flagged in these 16 samples
            const long long i = i0 + q * stride;
            const unsigned w[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if ((w[k >> 2] >> (8 * (k & 3))) & 0xffu) {
                    body[i * 16 + k] = E::nan();
                    ++mine;
                }
            }
        }
    }
    if (blockIdx.x == 0) {
        const long long tail0 = head + n_vec * 16;
        for (long long i = threadIdx.x; i < head + (n - tail0); i += blockDim.x) {
            const long long j = i < head ? i : tail0 + (i - head);
            if (flag[j]) {
                data[j] = E::nan();
                ++mine;
            }
        }
    }
    count_flagged(mine, n_flagged);
}

// Out of place, any alignment: out[i] = flag[i] ? NaN : data[i]
template <int KIND>
__global__ void __launch_bounds__(256) apply_flags_copy_kernel(const typename Elem<KIND>::type *__restrict__ data,
                                                               typename Elem<KIND>::type *__restrict__ out,
                                                               const unsigned char *__restrict__ flag, long long n,
                                                               unsigned long long *n_flagged)
{
    using E = Elem<KIND>;
    int mine = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // four independent loads in flight per thread
    for (; i + 3 * stride < n; i += 4 * stride) {
        typename E::type v[4];
        unsigned char f[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            f[k] = flag[i + k * stride];
            v[k] = data[i + k * stride];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            out[i + k * stride] = f[k] ? E::nan() : v[k];
            mine += f[k] != 0;
        }
    }
    for (; i < n; i += stride) {
        const unsigned char f = flag[i];
        out[i] = f ? E::nan() : data[i];
        mine += f != 0;
    }
    count_flagged(mine, n_flagged);
}

// Out of place, 16-byte vectors: a thread moves V = 16 / sizeof(element) consecutive elements per load (data, out
// 16-byte aligned, flag V-byte aligned), four loads in flight.  The NaN pattern is assembled per 32-bit word.
template <int KIND> __device__ __forceinline__ unsigned nan_word(int w)
{
    if (KIND == CNGI_ELEM_F32 || KIND == CNGI_ELEM_C64) return 0x7fc00000u;
    return (w & 1) ? 0x7ff80000u : 0u;   // little-endian halves of the float64 quiet NaN
}

template <int V> __device__ __forceinline__ unsigned load_flags(const unsigned char *p);
template <> __device__ __forceinline__ unsigned load_flags<1>(const unsigned char *p) { return *p; }
template <> __device__ __forceinline__ unsigned load_flags<2>(const unsigned char *p)
{
    return *reinterpret_cast<const unsigned short *>(p);
}
template <> __device__ __forceinline__ unsigned load_flags<4>(const unsigned char *p)
{
    return *reinterpret_cast<const unsigned *>(p);
}

template <int KIND>
__global__ void __launch_bounds__(256) apply_flags_copy_vec_kernel(const uint4 *__restrict__ data, uint4 *__restrict__ out,
                                                                   const unsigned char *__restrict__ flag, long long n,
                                                                   unsigned long long *n_flagged)
{
    using E = Elem<KIND>;
    using T = typename E::type;
    constexpr int WPE = sizeof(T) / 4;   // 32-bit words per element
    constexpr int V = 4 / WPE;           // elements per 16-byte vector
    const long long n_vec = n / V;
    int mine = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_vec; i0 += 4 * stride) {
        uint4 v[4];
        unsigned f[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long long i = i0 + q * stride;
            if (i < n_vec) {
                f[q] = load_flags<V>(flag + i * V);
                v[q] = data[i];
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long long i = i0 + q * stride;
            if (i >= n_vec) break;
            unsigned w[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if ((f[q] >> (8 * (k / WPE))) & 0xffu) w[k] = nan_word<KIND>(k);
#pragma unroll
            for (int e = 0; e < V; ++e) mine += ((f[q] >> (8 * e)) & 0xffu) != 0;
            out[i] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    if (blockIdx.x == 0) {               // the n % V trailing elements
        const T *ds = reinterpret_cast<const T *>(data);
        T *os = reinterpret_cast<T *>(out);
        for (long long j = n_vec * V + threadIdx.x; j < n; j += blockDim.x) {
            os[j] = flag[j] ? E::nan() : ds[j];
            mine += flag[j] != 0;
        }
    }
    count_flagged(mine, n_flagged);
}

template <int KIND>
int launch(const void *data, void *out, const unsigned char *flag, long long n, unsigned long long *n_flagged,
           cudaStream_t st)
{
    using T = typename Elem<KIND>::type;
    CNGI_REQUIRE(((uintptr_t)data % sizeof(T)) == 0 && ((uintptr_t)out % sizeof(T)) == 0,
                 "apply_flags: data / out are not aligned to the element size");
    const int sms = sm_count();
    if (out == data) {
        long long head = (long long)((16 - ((uintptr_t)flag & 15)) & 15);
        if (head > n) head = n;
        const long long n_vec = (n - head) / 16;
        const long long blocks = ceil_div(n_vec > 0 ? n_vec : 1, 256);
        const unsigned grid = (unsigned)(blocks < (long long)sms * 8 ? blocks : (long long)sms * 8);
        apply_flags_inplace_kernel<KIND><<<grid, 256, 0, st>>>((T *)out, flag, n, head, n_vec, n_flagged);
    } else {
        constexpr int V = 16 / (int)sizeof(T);
        const bool vec_ok = ((uintptr_t)data % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)flag % V) == 0;
        const long long blocks = ceil_div(vec_ok ? n / V + 1 : n, 256 * 4);
        const unsigned grid = (unsigned)(blocks < (long long)sms * 8 ? blocks : (long long)sms * 8);
        if (vec_ok)
            apply_flags_copy_vec_kernel<KIND><<<grid, 256, 0, st>>>((const uint4 *)data, (uint4 *)out, flag, n, n_flagged);
        else   // sliced views: element-wise
            apply_flags_copy_kernel<KIND><<<grid, 256, 0, st>>>((const T *)data, (T *)out, flag, n, n_flagged);
    }
    CNGI_CUDA_TRY(cudaGetLastError());
    return CNGI_OK;
}

}  // namespace
}  // namespace cngi

extern "C" int cngi_b200_apply_flags(const void *data, void *out, const uint8_t *flag, int64_t n_elem, int32_t elem_kind,
                                     uint64_t *n_flagged, void *stream)
{
    using namespace cngi;
    cudaStream_t st = (cudaStream_t)stream;
    CNGI_REQUIRE(n_elem >= 0, "apply_flags: n_elem < 0");
    if (n_elem == 0) return CNGI_OK;   // empty variable: nothing to do (pointers may be NULL)
    CNGI_REQUIRE(data && out && flag, "apply_flags: data, out and flag are required");
    if (out != data) {
        const size_t bytes = (size_t)n_elem * (elem_kind == CNGI_ELEM_F32 ? 4 : elem_kind == CNGI_ELEM_C128 ? 16 : 8);
        const char *a = (const char *)data, *b = (const char *)out;
        CNGI_REQUIRE(a + bytes <= b || b + bytes <= a, "apply_flags: out partially overlaps data");
    }
    unsigned long long *nf = reinterpret_cast<unsigned long long *>(n_flagged);
    switch (elem_kind) {
    case CNGI_ELEM_F32: return launch<CNGI_ELEM_F32>(data, out, flag, n_elem, nf, st);
    case CNGI_ELEM_F64: return launch<CNGI_ELEM_F64>(data, out, flag, n_elem, nf, st);
    case CNGI_ELEM_C64: return launch<CNGI_ELEM_C64>(data, out, flag, n_elem, nf, st);
    case CNGI_ELEM_C128: return launch<CNGI_ELEM_C128>(data, out, flag, n_elem, nf, st);
    default:
        set_error("apply_flags: elem_kind %d is not one of CNGI_ELEM_F32/F64/C64/C128 (integer variables would need the "
                  "reference's NaN -> int cast, which is undefined)", (int)elem_kind);
        return CNGI_ERR_UNSUPPORTED;
    }
}
