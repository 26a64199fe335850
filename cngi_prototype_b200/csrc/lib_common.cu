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
