// microbench.cu -- measures the ceiling that bounds the flush side of the gridders: the rate at which L2 accepts
// REDG.E.ADD.F32x2 reductions, counted in 32-byte sectors (the unit of ncu's
// l1tex__t_sectors_pipe_lsu_mem_global_op_red).  NVIDIA publishes no atomic peak, so the "atomic roofline" that
// BASELINE.json's metric asks for is measured on the box.  Not on the product path.
#include "common.cuh"

namespace cngi {

// pattern 0: every lane its own sector (scattered, the shape of a window row leaving with lanes along u);
// pattern 1: groups of 8 lanes write 8 consecutive cells (64 B: the shape of a window column leaving with lanes along v);
// pattern 2: a warp writes 32 consecutive cells (256 B).
__global__ void __launch_bounds__(256) red_rate_kernel(float2 *buf, unsigned long long n_cells, int pattern, int per_thread)
{
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    unsigned long long h = (tid >> 5) * 0x9E3779B97F4A7C15ULL + 0x632BE59BD9B4E019ULL;   // one stream per warp
    const float2 one = make_float2(1.0f, -1.0f);
    for (int i = 0; i < per_thread; ++i) {
        h ^= h >> 29, h *= 0xBF58476D1CE4E5B9ULL, h ^= h >> 32;
        unsigned long long cell;
        if (pattern == 0) {
            unsigned long long hl = (h + lane) * 0x94D049BB133111EBULL;
            hl ^= hl >> 31;
            cell = (hl % (n_cells / 4)) * 4;                       // one 8-byte cell in a sector of its own
        } else if (pattern == 1) {
            unsigned long long hg = (h + (lane >> 3)) * 0x94D049BB133111EBULL;
            hg ^= hg >> 31;
            cell = (hg % (n_cells / 8)) * 8 + (lane & 7);           // 8 lanes: 64 contiguous bytes
        } else {
            cell = (h % (n_cells / 32)) * 32 + lane;                // 32 lanes: 256 contiguous bytes
        }
        atomicAdd(buf + cell, one);   // REDG.E.ADD.F32x2
    }
}

// The design north_star sketches -- per-block shared-memory subgrids updated with fp atomics -- measured in isolation:
// every thread adds `per_thread` S x S stamps of (re, im) pairs into a TILE x TILE shared-memory subgrid at
// pseudo-random positions (no index math, no taps, no loads: an upper bound for that design), then the subgrid is
// flushed once with REDG.  On sm_100a atomicAdd(float) on shared memory is an ATOMS.CAST.SPIN compare-and-swap loop.
template <int TILE, int S> __global__ void __launch_bounds__(256) smem_atomic_rate_kernel(float2 *sink, int per_thread)
{
    __shared__ float2 tile[TILE * TILE];
    for (int i = threadIdx.x; i < TILE * TILE; i += blockDim.x) tile[i] = make_float2(0.f, 0.f);
    __syncthreads();
    unsigned h = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int i = 0; i < per_thread; ++i) {
        h = h * 1664525u + 1013904223u;
        const int u0 = (h >> 8) % (TILE - S + 1), v0 = (h >> 20) % (TILE - S + 1);
#pragma unroll
        for (int iu = 0; iu < S; ++iu)
#pragma unroll
            for (int iv = 0; iv < S; ++iv) {
                float2 *c = tile + (u0 + iu) * TILE + v0 + iv;
                atomicAdd(&c->x, 1.0f);
                atomicAdd(&c->y, -1.0f);
            }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TILE * TILE; i += blockDim.x) atomicAdd(sink + i, tile[i]);
}

}  // namespace cngi

// blocks x 256 threads, each adding per_thread 7x7 complex stamps into a 32x32 shared-memory subgrid with fp atomics;
// sink (>= 1024 float2) receives the flushed subgrids.  Tap updates issued: blocks*256*per_thread*49.
extern "C" int cngi_b200_microbench_smem_atomics(void *sink, int32_t blocks, int32_t per_thread, void *stream)
{
    using namespace cngi;
    CNGI_REQUIRE(sink && blocks > 0 && per_thread > 0, "microbench_smem_atomics: bad arguments");
    smem_atomic_rate_kernel<32, 7><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((float2 *)sink, per_thread);
    CNGI_CUDA_TRY(cudaGetLastError());
    return CNGI_OK;
}

// Launches `blocks` blocks of 256 threads, each thread issuing `per_thread` reductions into buf[0 .. n_cells).
// Sectors touched per warp instruction: 32 (pattern 0), 8 (pattern 1), 8 (pattern 2).
extern "C" int cngi_b200_microbench_red(void *buf, int64_t n_cells, int32_t pattern, int32_t blocks, int32_t per_thread,
                                        void *stream)
{
    using namespace cngi;
    CNGI_REQUIRE(buf && n_cells >= 64 && pattern >= 0 && pattern <= 2 && blocks > 0 && per_thread > 0,
                 "microbench_red: bad arguments");
    red_rate_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((float2 *)buf, (unsigned long long)n_cells, pattern,
                                                                        per_thread);
    CNGI_CUDA_TRY(cudaGetLastError());
    return CNGI_OK;
}
