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
// handled one by one by the first block.
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
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) {
        const uint4 f = __ldg(fv + i);
        if ((f.x | f.y | f.z | f.w) == 0u) continue;   // the common case: nothing flagged in these 16 samples
        const unsigned w[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if ((w[k >> 2] >> (8 * (k & 3))) & 0xffu) {
                body[i * 16 + k] = E::nan();
                ++mine;
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

// Out of place: out[i] = flag[i] ? NaN : data[i]
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
        const long long blocks = ceil_div(n, 256 * 4);
        const unsigned grid = (unsigned)(blocks < (long long)sms * 8 ? blocks : (long long)sms * 8);
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
