// multimem.cu -- in-switch (NVLS) reductions of uv-grids over a multicast mapping of symmetric memory.
//
// The time-sharded continuum step sums N partial uv-grids (268 MB complex64 at 4096^2 x 2 pol) onto the rank that runs
// the FFT, and N partial density planes (134 MB fp64) onto every rank (distributed.py; the role of the reference's dask
// tree sum, _standard_grid.py:109-120).  NCCL does both with 24-32 blocks of 640 threads on EVERY rank, taken from a
// gridder that is running at the same time.  With the buffers in symmetric memory mapped through an NVSwitch multicast
// address, a `multimem.ld_reduce` returns the sum over all ranks computed inside the switch:
//   * reduce-to-root: ONLY the root runs a kernel (a few blocks streaming the multicast range into its own buffer); the
//     other ranks' SMs are not involved at all, their HBM is read over NVLink;
//   * all-reduce: every rank reduces its 1/N slice with multimem.ld_reduce and broadcasts it with multimem.st.
// The cross-rank ordering (all partial grids complete before anybody reads them; nobody overwrites a buffer that is still
// being read) is the caller's: distributed.SymmetricCollectives brackets the kernels with the symmetric-memory barrier.
#include "common.cuh"
#include <algorithm>

namespace cngi {

// A multimem.ld_reduce is a round trip through the switch (microseconds): bandwidth comes from bytes in flight, so every
// thread keeps U independent loads outstanding (32 blocks x 512 threads x 8 x 16 B = 2 MB).
constexpr int kMmUnroll = 8;

__global__ void __launch_bounds__(512) mm_reduce_f32_kernel(const float *mc, float *dst, long long n4)   // n4 = float4 count
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (kMmUnroll - 1) * stride < n4; i += kMmUnroll * stride) {
        float4 v[kMmUnroll];
#pragma unroll
        for (int u = 0; u < kMmUnroll; ++u)
            asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                         : "l"(mc + 4 * (i + u * stride))
                         : "memory");
#pragma unroll
        for (int u = 0; u < kMmUnroll; ++u) reinterpret_cast<float4 *>(dst)[i + u * stride] = v[u];
    }
    for (; i < n4; i += stride) {
        float4 v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "l"(mc + 4 * i)
                     : "memory");
        reinterpret_cast<float4 *>(dst)[i] = v;
    }
}

// all-reduce of this rank's slice [lo, hi) (in doubles): sum through the switch, broadcast through the switch
__global__ void __launch_bounds__(512) mm_allreduce_f64_kernel(double *mc, long long lo, long long hi)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (kMmUnroll - 1) * stride < hi; i += kMmUnroll * stride) {
        double v[kMmUnroll];
#pragma unroll
        for (int u = 0; u < kMmUnroll; ++u)
            asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(v[u]) : "l"(mc + i + u * stride) : "memory");
#pragma unroll
        for (int u = 0; u < kMmUnroll; ++u)
            asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc + i + u * stride), "d"(v[u]) : "memory");
    }
    for (; i < hi; i += stride) {
        double v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(v) : "l"(mc + i) : "memory");
        asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc + i), "d"(v) : "memory");
    }
}

}  // namespace cngi

extern "C" int cngi_b200_multimem_reduce_f32(const void *multicast_ptr, void *dst, int64_t n_floats, int32_t n_blocks, void *stream)
{
    using namespace cngi;
    CNGI_REQUIRE(multicast_ptr && dst, "multimem_reduce_f32: null pointer");
    CNGI_REQUIRE(n_floats >= 0 && n_floats % 4 == 0 && ((uintptr_t)multicast_ptr % 16) == 0 && ((uintptr_t)dst % 16) == 0,
                 "multimem_reduce_f32: the range must be 16-byte aligned and a multiple of 4 floats");
    if (n_floats == 0) return CNGI_OK;
    if (n_blocks <= 0) n_blocks = 32;
    mm_reduce_f32_kernel<<<(unsigned)n_blocks, 512, 0, (cudaStream_t)stream>>>((const float *)multicast_ptr, (float *)dst, n_floats / 4);
    CNGI_CUDA_TRY(cudaGetLastError());
    return CNGI_OK;
}

extern "C" int cngi_b200_multimem_allreduce_f64(void *multicast_ptr, int64_t n_doubles, int32_t rank, int32_t world_size,
                                                int32_t n_blocks, void *stream)
{
    using namespace cngi;
    CNGI_REQUIRE(multicast_ptr && world_size > 0 && rank >= 0 && rank < world_size, "multimem_allreduce_f64: bad arguments");
    CNGI_REQUIRE(((uintptr_t)multicast_ptr % 8) == 0 && n_doubles >= 0, "multimem_allreduce_f64: misaligned range");
    const int64_t per = (n_doubles + world_size - 1) / world_size;
    const int64_t lo = std::min<int64_t>(n_doubles, per * rank), hi = std::min<int64_t>(n_doubles, lo + per);
    if (hi <= lo) return CNGI_OK;
    if (n_blocks <= 0) n_blocks = 32;
    mm_allreduce_f64_kernel<<<(unsigned)n_blocks, 512, 0, (cudaStream_t)stream>>>((double *)multicast_ptr, lo, hi);
    CNGI_CUDA_TRY(cudaGetLastError());
    return CNGI_OK;
}
