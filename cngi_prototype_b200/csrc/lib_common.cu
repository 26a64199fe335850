// lib_common.cu -- error plumbing, device checks (C ABI housekeeping).
#include "common.cuh"
#include <mutex>

namespace cngi {

static thread_local char g_error[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int sm_count()
{
    static int cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// Stream-ordered scratch (cudaMallocAsync) is used by several entry points; with the default release threshold (0)
// the pool hands its memory back to the driver at every synchronisation and each call pays a real cudaMalloc
// (measured: 0.45 ms per cngi_b200_direction_rotate call).  Raised once per device.
int tune_pool_once()
{
    static bool done[64];
    int dev = 0;
    CNGI_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !done[dev]) {
        cudaMemPool_t pool;
        CNGI_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t keep = UINT64_MAX;
        CNGI_CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        done[dev] = true;
    }
    return CNGI_OK;
}

__global__ void uv_scale_kernel(const double *freq, int n_chan, double dl, double dm, int n_u, int n_v, double *table)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chan) return;
    table[c] = uv_scale_of(freq[c], dl, n_u);
    table[n_chan + c] = uv_scale_of(freq[c], dm, n_v);
}

int make_uv_scale_table(const double *freq, int n_chan, double dl, double dm, int n_u, int n_v, cudaStream_t st,
                        double **table)
{
    if (int rc = tune_pool_once()) return rc;
    CNGI_CUDA_TRY(cudaMallocAsync((void **)table, (size_t)2 * n_chan * sizeof(double), st));
    uv_scale_kernel<<<(n_chan + 127) / 128, 128, 0, st>>>(freq, n_chan, dl, dm, n_u, n_v, *table);
    CNGI_CUDA_TRY(cudaGetLastError());
    return CNGI_OK;
}

}  // namespace cngi

extern "C" int cngi_b200_abi_version(void) { return CNGI_B200_ABI_VERSION; }

extern "C" const char *cngi_b200_last_error(void) { return cngi::g_error; }

extern "C" int cngi_b200_check_device(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cngi::set_error("no CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return CNGI_ERR_NO_DEVICE;
    }
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        cngi::set_error("cannot query the current CUDA device");
        return CNGI_ERR_NO_DEVICE;
    }
    if (major != 10) {
        cngi::set_error("libcngi_b200 is built for sm_100a only; current device has compute capability %d.x", major);
        return CNGI_ERR_NO_DEVICE;
    }
    return CNGI_OK;
}
