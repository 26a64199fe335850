// gcf.cu -- N2: A-term gridding convolution functions on the device.
// Replaces the a_term branch of ngcasa/imaging/make_gridding_convolution_function.py:161-311:
//   make_baseline_patterns :394-412 with _casa_airy_disk_rorder / _airy_disk_rorder (_make_pb_symmetric.py:187-235,
//   :135-183)  ->  gcf_pattern_kernel      (voltage pattern product A_i A_j and its PB^2 counterpart, written
//                                           already ifftshift-ed)
//   real(fftshift(fft2(ifftshift(.)))) :246-247  ->  cuFFT Z2Z forward, batch 2; the fftshift is an index remap in
//                                           the consumers, never a pass over the data
//   calc_conv_size :414-457                 ->  gcf_stats_kernel (max / min |.| of the weight kernel) +
//                                           the parallel threshold walk at the head of gcf_finalize_kernel
//   resize_and_calc_support :361-392        ->  gcf_finalize_kernel (crop about the centre, window sum, normalise)
//   make_phase_gradient :331-359            ->  gcf_phase_gradient_kernel (the SIN world2pix is 2 numbers per field,
//                                           done by the host mirror)
// One (antenna-type pair, PB frequency) item = two n_pad^2 complex planes in a reusable workspace; per item the data
// is written once, transformed in place, and read once for the max/min plus a conv_size^2 window.
#include "common.cuh"
#include <cufft.h>
#include <mutex>

namespace cngi {

struct DishParams {
    double aperture;      // dish / 2
    double a, b;          // casa: a = area_ratio, b = length_ratio; airy: a = e = blockage / dish
    int kind;             // 0: no blockage, 1: casa_airy, 2: airy
};

__device__ __forceinline__ double voltage(const DishParams &d, double rad_k)
{
    const double r = rad_k * d.aperture;
    const double first = j1(r);
    if (d.kind == 0) return 2.0 * first / r;
    if (d.kind == 1) {   // _make_pb_symmetric.py:227-229
        const double rl = r * d.b;
        return (d.a * 2.0 * first / r - 2.0 * j1(rl) / rl) / (d.a - 1.0);
    }
    return (2.0 * first / r - 2.0 * d.a * j1(r * d.a) / r) / (1.0 - d.a * d.a);   // :175-176
}

// planes[0] = A_i A_j, planes[1] = A_i^2 A_j^2 (ipower 1 / 2, :180-190), complex with zero imaginary part,
// stored at the ifftshift-ed position so that cuFFT sees ifftshift(pattern).
__global__ void __launch_bounds__(256)
gcf_pattern_kernel(double2 *__restrict__ planes, int n0, int n1, double cell0, double cell1, double k, DishParams di,
                   DishParams dj, bool same)
{
    const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    const int i0 = blockIdx.y;
    if (i1 >= n1) return;
    const int c0 = n0 / 2, c1 = n1 / 2;
    double vi, vj;
    if (i0 == c0 && i1 == c1) {
        vi = vj = 1.0;       // centre fixed to 1 (:179,:231)
    } else {
        const double x = (double)(i0 - c0) * cell0, y = (double)(i1 - c1) * cell1;
        const double rad_k = sqrt(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y))) * k;
        vi = voltage(di, rad_k);
        vj = same ? vi : voltage(dj, rad_k);
    }
    // ifftshift: out[j] = in[(j + n/2) % n]  <=>  in index i lands at j = (i - n/2) mod n
    const int j0 = i0 >= c0 ? i0 - c0 : i0 - c0 + n0;
    const int j1_ = i1 >= c1 ? i1 - c1 : i1 - c1 + n1;
    const size_t o = (size_t)j0 * n1 + j1_;
    planes[o] = make_double2(vi * vj, 0.0);
    planes[(size_t)n0 * n1 + o] = make_double2((vi * vi) * (vj * vj), 0.0);
}

// stats[0] = max |Re F|, stats[1] = min |Re F| over the weight plane (bit patterns of non-negative doubles order
// like unsigned integers).
__global__ void __launch_bounds__(256) gcf_stats_kernel(const double2 *__restrict__ w, long long n,
                                                        unsigned long long *stats)
{
    double mx = 0.0, mn = INFINITY;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double a = fabs(w[i].x);
        mx = fmax(mx, a);
        mn = fmin(mn, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(stats, (unsigned long long)__double_as_longlong(mx));
        atomicMin(stats + 1, (unsigned long long)__double_as_longlong(mn));
    }
}

struct FinalizeParams {
    const double2 *planes;     // [2, n0, n1] unshifted spectra
    const unsigned long long *stats;
    double *conv_kernel, *weight_conv_kernel;   // this item's [cu, cv] planes
    long long *support;        // this item's [2]
    int *status;
    double cut_level;
    int n0, n1, cu, cv, os0, os1, max0, max1;
};

// shifted[k0, k1] = F[(k0 - h0) mod n0, (k1 - h1) mod n1]
__device__ __forceinline__ double shifted_re(const double2 *pl, int k0, int k1, int n0, int n1)
{
    int f0 = k0 - n0 / 2, f1 = k1 - n1 / 2;
    if (f0 < 0) f0 += n0;
    if (f1 < 0) f1 += n1;
    return pl[(size_t)f0 * n1 + f1].x;
}

__device__ double block_sum(double v, double *sh)
{
    const int tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (tid < o) sh[tid] += sh[tid + o];
        __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(256) gcf_finalize_kernel(FinalizeParams p)
{
    __shared__ double sh[256];
    __shared__ int first[2];
    const int tid = threadIdx.x;
    const double2 *pb = p.planes, *wt = p.planes + (size_t)p.n0 * p.n1;
    const double mx = __longlong_as_double((long long)p.stats[0]), mn = __longlong_as_double((long long)p.stats[1]);
    const double cut = p.cut_level * mx;
    const int h0 = p.n0 / 2, h1 = p.n1 / 2;
    if (tid == 0) {
        first[0] = p.n0;
        first[1] = p.n1;
        if (!(mn < cut)) atomicOr(p.status, 1);      // assert at :423
    }
    __syncthreads();
    // calc_conv_size :426-441: first index >= centre where the (signed) weight kernel is <= cut, along +x and +y
    for (int i = h0 + tid; i < p.n0; i += blockDim.x)
        if (!(shifted_re(wt, i, h1, p.n0, p.n1) > cut)) {
            atomicMin(&first[0], i);
            break;
        }
    for (int i = h1 + tid; i < p.n1; i += blockDim.x)
        if (!(shifted_re(wt, h0, i, p.n0, p.n1) > cut)) {
            atomicMin(&first[1], i);
            break;
        }
    __syncthreads();
    if (first[0] >= p.n0 || first[1] >= p.n1) {      // asserts at :429,:440
        if (tid == 0) atomicOr(p.status, 2);
        return;
    }
    const int sx = (__double2int_rz(0.5 + (double)(first[0] - h0) / (double)p.os0) + 1) * 2 + 1;
    const int sy = (__double2int_rz(0.5 + (double)(first[1] - h1) / (double)p.os1) + 1) * 2 + 1;
    if (tid == 0 && !(sx < p.max0 && sy < p.max1)) atomicOr(p.status, 4);   // :447-448
    const int s = max(sx, sy);
    if (tid == 0) p.support[0] = p.support[1] = s;
    // resize_and_calc_support :376-387
    const int st0 = h0 - p.cu / 2, st1 = h1 - p.cv / 2;
    const int em0 = (s + 1) * p.os0, em1 = (s + 1) * p.os1;
    const int e0 = p.cu / 2 - em0 / 2, e1 = p.cv / 2 - em1 / 2;
    double acc_pb = 0.0, acc_wt = 0.0;
    for (int q = tid; q < em0 * em1; q += blockDim.x) {
        const int a = e0 + q / em1, b = e1 + q % em1;
        if (a >= 0 && a < p.cu && b >= 0 && b < p.cv) {
            acc_pb += shifted_re(pb, st0 + a, st1 + b, p.n0, p.n1);
            acc_wt += shifted_re(wt, st0 + a, st1 + b, p.n0, p.n1);
        }
    }
    const double norm_pb = block_sum(acc_pb, sh) / (double)(p.os0 * p.os1);
    const double norm_wt = block_sum(acc_wt, sh) / (double)(p.os0 * p.os1);
    for (int q = tid; q < p.cu * p.cv; q += blockDim.x) {
        const int a = q / p.cv, b = q % p.cv;
        p.conv_kernel[q] = shifted_re(pb, st0 + a, st1 + b, p.n0, p.n1) / norm_pb;
        p.weight_conv_kernel[q] = shifted_re(wt, st0 + a, st1 + b, p.n0, p.n1) / norm_wt;
    }
}

// exp(i (x pix0 + y pix1)), x = i - cu//2, y = j - cv//2   (:351-358)
__global__ void gcf_phase_gradient_kernel(const double *__restrict__ pix, double2 *__restrict__ out, int cu, int cv)
{
    const int f = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= cu * cv) return;
    const int a = q / cv, b = q % cv;
    const double arg = __dadd_rn(__dmul_rn((double)(a - cu / 2), pix[2 * f]), __dmul_rn((double)(b - cv / 2), pix[2 * f + 1]));
    double s, c;
    sincos(arg, &s, &c);
    out[(size_t)f * cu * cv + q] = make_double2(c, s);
}


// ---- make_pb: primary-beam images pb[l, m, chan, pol, dish] = voltage^ipower  (make_pb.py:95-118 with _airy_disk /
// _casa_airy_disk, _make_pb_symmetric.py:26-132).  One thread per (pixel, chan): consecutive threads write consecutive
// n_pol * n_dish runs, so the stores are coalesced.
struct PbParams {
    double *pb;
    const double *freq;
    long long n_items;       // n_l * n_m * n_chan
    int n_l, n_m, n_chan, n_pol, n_dish, c0, c1, ipower;
    double cell0, cell1;
    DishParams dish[8];
};

__global__ void __launch_bounds__(256) pb_kernel(PbParams p)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_items) return;
    const int c = (int)(i % p.n_chan);
    const long long pix = i / p.n_chan;
    const int m = (int)(pix % p.n_m), l = (int)(pix / p.n_m);
    const bool centre = (l == p.c0 && m == p.c1);
    const double x = (double)(l - p.c0) * p.cell0, y = (double)(m - p.c1) * p.cell1;
    const double k = __ddiv_rn(__dmul_rn(6.283185307179586, p.freq[c]), kSpeedOfLight);
    const double rad_k = sqrt(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y))) * k;
    double *out = p.pb + i * p.n_pol * p.n_dish;
    for (int d = 0; d < p.n_dish; ++d) {
        double v = 1.0;
        if (!centre) {
            v = voltage(p.dish[d], rad_k);
            if (p.ipower == 2) v = v * v;
        }
        for (int q = 0; q < p.n_pol; ++q) out[q * p.n_dish + d] = v;
    }
}

// The Z2Z plan (batch 2, n_pad^2) of the last geometry is kept: creating it costs more than one item's kernels.
// One plan cannot be driven from two host threads at once, so cngi_b200_make_gcf calls are serialised on this mutex.
static std::mutex g_gcf_mutex;
static struct {
    bool valid = false;
    int dev = -1;
    long long n0 = 0, n1 = 0;
    cufftHandle plan = 0;
    cudaEvent_t done = nullptr;   // end of the last call's work: the next call's stream waits for it before reusing the plan
} g_gcf_plan;

static DishParams dish_params(int function, double dish, double blockage)
{
    DishParams d;
    d.aperture = dish / 2;
    if (blockage == 0.0) {
        d.kind = 0;
        d.a = d.b = 0.0;
    } else if (function == CNGI_PB_CASA_AIRY) {
        d.kind = 1;
        d.a = (dish / blockage) * (dish / blockage);
        d.b = dish / blockage;
    } else {
        d.kind = 2;
        d.a = blockage / dish;
        d.b = 0.0;
    }
    return d;
}

}  // namespace cngi

extern "C" int cngi_b200_make_gcf(const cngi_gcf_args *a, void *stream)
{
    using namespace cngi;
    cudaStream_t st = (cudaStream_t)stream;
    CNGI_REQUIRE(a != nullptr, "make_gcf: args is NULL");
    CNGI_REQUIRE(a->function == CNGI_PB_AIRY || a->function == CNGI_PB_CASA_AIRY, "make_gcf: unknown function %d", a->function);
    const long long n0 = a->n_pad[0], n1 = a->n_pad[1], cu = a->conv_size[0], cv = a->conv_size[1];
    CNGI_REQUIRE(n0 > 0 && n1 > 0 && n0 < 65536 && n1 < (1 << 30), "make_gcf: bad padded size");
    CNGI_REQUIRE(cu > 0 && cv > 0 && cu <= n0 && cv <= n1, "make_gcf: conv_size must fit inside the padded image");
    CNGI_REQUIRE(a->oversampling[0] > 0 && a->oversampling[1] > 0, "make_gcf: bad oversampling");
    CNGI_REQUIRE(a->n_dish > 0 && a->n_pair > 0 && a->n_freq > 0, "make_gcf: empty dish / pair / frequency list");
    CNGI_REQUIRE(a->dish_diameter_host && a->blockage_diameter_host && a->ant_pairs_host && a->pb_freq_host,
                 "make_gcf: host parameter arrays are required");
    CNGI_REQUIRE(a->conv_kernel && a->weight_conv_kernel && a->support && a->status, "make_gcf: outputs are required");
    for (long long k = 0; k < a->n_pair; ++k)
        CNGI_REQUIRE(a->ant_pairs_host[2 * k] >= 0 && a->ant_pairs_host[2 * k] < a->n_dish &&
                     a->ant_pairs_host[2 * k + 1] >= 0 && a->ant_pairs_host[2 * k + 1] < a->n_dish,
                     "make_gcf: antenna-type pair %lld out of range", k);

    if (int rc = tune_pool_once()) return rc;
    std::lock_guard<std::mutex> lock(g_gcf_mutex);
    int dev = 0;
    CNGI_CUDA_TRY(cudaGetDevice(&dev));
    if (!(g_gcf_plan.valid && g_gcf_plan.dev == dev && g_gcf_plan.n0 == n0 && g_gcf_plan.n1 == n1)) {
        if (g_gcf_plan.valid) cufftDestroy(g_gcf_plan.plan);
        g_gcf_plan.valid = false;
        int dims[2] = {(int)n0, (int)n1};
        cufftResult fr = cufftPlanMany(&g_gcf_plan.plan, 2, dims, nullptr, 1, (int)(n0 * n1), nullptr, 1, (int)(n0 * n1),
                                       CUFFT_Z2Z, 2);
        if (fr != CUFFT_SUCCESS) {
            set_error("make_gcf: cufftPlanMany(%lld x %lld, batch 2) failed with %d", n0, n1, (int)fr);
            return CNGI_ERR_CUDA;
        }
        g_gcf_plan.valid = true, g_gcf_plan.dev = dev, g_gcf_plan.n0 = n0, g_gcf_plan.n1 = n1;
        if (g_gcf_plan.done) cudaEventDestroy(g_gcf_plan.done);
        g_gcf_plan.done = nullptr;
        CNGI_CUDA_TRY(cudaEventCreateWithFlags(&g_gcf_plan.done, cudaEventDisableTiming));
    } else {
        CNGI_CUDA_TRY(cudaStreamWaitEvent(st, g_gcf_plan.done, 0));
    }
    const cufftHandle plan = g_gcf_plan.plan;
    double2 *planes = nullptr;
    unsigned long long *stats = nullptr;
    int rc = CNGI_OK;
    cudaError_t e = cudaMallocAsync((void **)&planes, (size_t)2 * n0 * n1 * sizeof(double2), st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&stats, (size_t)2 * a->n_pair * a->n_freq * sizeof(unsigned long long), st);
    if (e == cudaSuccess && cufftSetStream(plan, st) != CUFFT_SUCCESS) {
        set_error("make_gcf: cufftSetStream failed");
        rc = CNGI_ERR_CUDA;
    }
    const dim3 pgrid((unsigned)ceil_div(n1, 256), (unsigned)n0);
    const int sblocks = (int)std::min<long long>(ceil_div(n0 * n1, 256), (long long)sm_count() * 8);
    for (long long k = 0; e == cudaSuccess && rc == CNGI_OK && k < a->n_pair; ++k) {
        const long long di = a->ant_pairs_host[2 * k], dj = a->ant_pairs_host[2 * k + 1];
        const DishParams pi = dish_params(a->function, a->dish_diameter_host[di], a->blockage_diameter_host[di]);
        const DishParams pj = dish_params(a->function, a->dish_diameter_host[dj], a->blockage_diameter_host[dj]);
        for (long long q = 0; q < a->n_freq; ++q) {
            const long long item = k * a->n_freq + q;
            const double kwave = (6.283185307179586 * a->pb_freq_host[q]) / kSpeedOfLight;   // (2 pi f) / c
            unsigned long long *sti = stats + 2 * item;
            e = cudaMemsetAsync(sti, 0, sizeof(unsigned long long), st);                 // max := +0.0
            if (e == cudaSuccess) e = cudaMemsetAsync(sti + 1, 0x7f, sizeof(unsigned long long), st);   // min := huge
            if (e != cudaSuccess) break;
            gcf_pattern_kernel<<<pgrid, 256, 0, st>>>(planes, (int)n0, (int)n1, a->pb_cell[0], a->pb_cell[1], kwave, pi,
                                                      pj, di == dj);
            if (cufftExecZ2Z(plan, (cufftDoubleComplex *)planes, (cufftDoubleComplex *)planes, CUFFT_FORWARD) !=
                CUFFT_SUCCESS) {
                set_error("make_gcf: cufftExecZ2Z failed");
                rc = CNGI_ERR_CUDA;
                break;
            }
            gcf_stats_kernel<<<sblocks, 256, 0, st>>>(planes + (size_t)n0 * n1, n0 * n1, sti);
            FinalizeParams fp;
            fp.planes = planes;
            fp.stats = sti;
            fp.conv_kernel = a->conv_kernel + (size_t)item * cu * cv;
            fp.weight_conv_kernel = a->weight_conv_kernel + (size_t)item * cu * cv;
            fp.support = (long long *)a->support + 2 * item;
            fp.status = a->status;
            fp.cut_level = a->support_cut_level;
            fp.n0 = (int)n0, fp.n1 = (int)n1, fp.cu = (int)cu, fp.cv = (int)cv;
            fp.os0 = a->oversampling[0], fp.os1 = a->oversampling[1];
            fp.max0 = a->max_support[0], fp.max1 = a->max_support[1];
            gcf_finalize_kernel<<<1, 256, 0, st>>>(fp);
            e = cudaGetLastError();
            if (e != cudaSuccess) break;
        }
    }
    if (planes) cudaFreeAsync(planes, st);
    if (stats) cudaFreeAsync(stats, st);
    cudaEventRecord(g_gcf_plan.done, st);
    if (rc != CNGI_OK) return rc;
    CNGI_CUDA_TRY(e);
    return CNGI_OK;
}

extern "C" int cngi_b200_phase_gradient(const double *pix, int64_t n_field, int64_t cu, int64_t cv, void *phase_gradient,
                                        void *stream)
{
    using namespace cngi;
    CNGI_REQUIRE(pix && phase_gradient && n_field >= 0 && cu > 0 && cv > 0 && n_field < 65536 && cu * cv < (1LL << 31),
                 "phase_gradient: bad arguments");
    if (n_field == 0) return CNGI_OK;
    const dim3 grid((unsigned)ceil_div(cu * cv, 256), (unsigned)n_field);
    gcf_phase_gradient_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pix, (double2 *)phase_gradient, (int)cu, (int)cv);
    CNGI_CUDA_TRY(cudaGetLastError());
    return CNGI_OK;
}

extern "C" int cngi_b200_make_pb(const cngi_pb_args *a, void *stream)
{
    using namespace cngi;
    cudaStream_t st = (cudaStream_t)stream;
    CNGI_REQUIRE(a != nullptr && a->pb != nullptr, "make_pb: null args or output");
    CNGI_REQUIRE(a->function == CNGI_PB_AIRY || a->function == CNGI_PB_CASA_AIRY, "make_pb: unknown function %d", a->function);
    CNGI_REQUIRE(a->ipower == 1 || a->ipower == 2, "make_pb: ipower must be 1 (voltage pattern) or 2 (primary beam)");
    CNGI_REQUIRE(a->image_size[0] > 0 && a->image_size[1] > 0 && a->n_chan > 0 && a->n_pol > 0, "make_pb: empty image");
    CNGI_REQUIRE(a->n_dish >= 1 && a->n_dish <= 8, "make_pb: 1..8 dish types");
    CNGI_REQUIRE(a->freq_chan_host && a->dish_diameter_host && a->blockage_diameter_host, "make_pb: host parameter arrays are required");
    CNGI_REQUIRE(a->image_size[0] < (1LL << 31) && a->image_size[1] < (1LL << 31) && a->n_chan < (1LL << 31), "make_pb: axis too long");
    if (int rc = tune_pool_once()) return rc;
    PbParams p{};
    p.pb = a->pb;
    p.n_l = (int)a->image_size[0], p.n_m = (int)a->image_size[1], p.n_chan = (int)a->n_chan, p.n_pol = (int)a->n_pol;
    p.n_dish = (int)a->n_dish, p.c0 = (int)a->image_center[0], p.c1 = (int)a->image_center[1], p.ipower = a->ipower;
    p.cell0 = a->cell_size[0], p.cell1 = a->cell_size[1];
    p.n_items = (long long)p.n_l * p.n_m * p.n_chan;
    for (int d = 0; d < p.n_dish; ++d) p.dish[d] = dish_params(a->function, a->dish_diameter_host[d], a->blockage_diameter_host[d]);
    double *freq = nullptr;
    CNGI_CUDA_TRY(cudaMallocAsync((void **)&freq, (size_t)p.n_chan * sizeof(double), st));
    cudaError_t e = cudaMemcpyAsync(freq, a->freq_chan_host, (size_t)p.n_chan * sizeof(double), cudaMemcpyHostToDevice, st);
    p.freq = freq;
    const long long blocks = ceil_div(p.n_items, 256);
    if (e == cudaSuccess && blocks < (1LL << 31)) {
        pb_kernel<<<(unsigned)blocks, 256, 0, st>>>(p);
        e = cudaGetLastError();
    }
    cudaFreeAsync(freq, st);
    CNGI_REQUIRE(blocks < (1LL << 31), "make_pb: image cube too large for one launch");
    CNGI_CUDA_TRY(e);
    return CNGI_OK;
}
