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
