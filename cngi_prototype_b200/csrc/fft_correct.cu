// fft_correct.cu -- A9/A10 of SURVEY.md section 8: uv-grid -> corrected image.
//   fftshift(ifft2(ifftshift(G))) -> centre crop -> .real * (n_u*n_v)     make_image.py:116-120
//   (img / sum_weight[0 -> 1]) / correcting image                          make_image.py:123-130
//   ... / (sinc x sinc * PS_CORR_IMAGE * PB), pb_limit mask, f32 round trip _normalize.py:39-89
//
// cuFFT does the transform (unnormalised inverse == numpy's ifft2 times n_u*n_v, so the reference's
// "* (n_u*n_v)" is free).  Neither shift moves data:
//   ifftshift on the INPUT  is a phase ramp exp(-2 pi i h m / n), h = n//2, on the OUTPUT (== (-1)^m for
//   even n; tabulated per axis for odd n, where fftshift != ifftshift), and
//   fftshift  on the OUTPUT is an index remap folded, together with the crop, into the one pass of the
//   post kernel, which reads only the cropped window and writes the real, normalised image.
#include "common.cuh"
#include "fft_bluestein.cuh"
#include <cufft.h>
#include <cstdlib>
#include <vector>
#include <algorithm>

struct cngi_fft_plan {
    int64_t n_u, n_v, max_planes;
    int32_t precision;
    cufftHandle plan;          // batch = max_planes
    cufftHandle plan_tail;     // lazily created for the last, smaller batch
    int64_t tail_planes;
    void *work;                // complex [max_planes, n_u, n_v]
    double2 *phase_u, *phase_v;   // exp(-2 pi i h m / n) per axis
    // complex64 grids whose sides are n1 * (a prime cuFFT has no radix for) -- 4915, 9830, ... from the reference's default
    // fft_padding of 1.2: shared-memory Bluestein passes (fft_bluestein.cu) instead of cuFFT; [0] inverse, [1] forward
    bool use_blu;
    cngi::BluAxis blu_u[2], blu_v[2];
    bool blu_ready[2];
};

namespace cngi {

template <typename T> __global__ void real_to_complex_kernel(const T *in, typename Cplx<T>::type *out, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        typename Cplx<T>::type c;
        c.x = in[i];
        c.y = (T)0;
        out[i] = c;
    }
}

struct PostParams {
    const void *spec;          // complex [planes, n_u, n_v]: unnormalised inverse DFT of the unshifted grid
    void *image;               // real [planes, n_l, n_m]
    const double2 *phase_u, *phase_v;
    const double *sum_weight, *corr_u, *corr_v;
    const void *norm_image, *pb_image;
    long long norm_planes, pb_planes;
    double pb_limit;
    int n_u, n_v, n_l, n_m, start_u, start_v;
    int plane0;                // first global plane of this batch
    int roundtrip;
};

// everything after the spectrum value z of image pixel (l, m) of global plane gp has been fetched
template <typename T>
__device__ __forceinline__ void post_pixel(const PostParams &p, typename Cplx<T>::type z, int l, int m, int mu, int mv, int gp)
{
    const double2 pu = p.phase_u[mu], pv = p.phase_v[mv];
    const double pr = pu.x * pv.x - pu.y * pv.y, pi = pu.x * pv.y + pu.y * pv.x;
    double val = (double)z.x * pr - (double)z.y * pi;       // Re(phase * z)
    if (p.sum_weight) {
        double sw = p.sum_weight[gp];
        if (sw == 0.0) sw = 1.0;
        val = val / sw;
    }
    const long long pix = (long long)l * p.n_m + m;
    const long long npix = (long long)p.n_l * p.n_m;
    double div = 1.0;
    bool has_div = false;
    if (p.corr_u) {
        div = p.corr_u[l] * p.corr_v[m];
        has_div = true;
    }
    if (p.norm_image) {
        const double nv = (double)((const T *)p.norm_image)[(p.norm_planes == 1 ? 0 : gp) * npix + pix];
        div = has_div ? div * nv : nv;
        has_div = true;
    }
    if (has_div) val = val / div;
    if (p.pb_image) {
        const double pb = (double)((const T *)p.pb_image)[(p.pb_planes == 1 ? 0 : gp) * npix + pix];
        if (pb < p.pb_limit) val = 0.0;
    }
    if (p.roundtrip) val = (double)(float)val;
    ((T *)p.image)[(long long)gp * npix + pix] = (T)val;
}

// fftshift + crop: image pixel l (m) is shifted-spectrum index k = start + l, i.e. DFT bin (k - h) mod n
__device__ __forceinline__ int post_bin(int start, int i, int n)
{
    int b = start + i - n / 2;
    return b < 0 ? b + n : b;
}

template <typename T> __global__ void __launch_bounds__(256) fft_post_kernel(PostParams p)
{
    using CT = typename Cplx<T>::type;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;   // fastest image axis
    const int l = blockIdx.y;
    const int pl = blockIdx.z;                              // plane within the batch
    if (m >= p.n_m) return;
    const int mu = post_bin(p.start_u, l, p.n_u), mv = post_bin(p.start_v, m, p.n_v);
    const CT z = ((const CT *)p.spec)[((long long)pl * p.n_u + mu) * p.n_v + mv];
    post_pixel<T>(p, z, l, m, mu, mv, p.plane0 + pl);
}

// The same pass over a spectrum stored TRANSPOSED (spec[plane][v bin][u bin], what the shared-memory Bluestein passes
// leave: rows out transposed, then rows again): 32 x 32 tiles go through shared memory so that both the reads (along u)
// and the image writes (along m) are coalesced.  blockDim = (32, 8).
template <typename T> __global__ void __launch_bounds__(256) fft_post_transposed_kernel(PostParams p)
{
    using CT = typename Cplx<T>::type;
    __shared__ CT tile[32][33];
    const int m0 = blockIdx.x * 32, l0 = blockIdx.y * 32, pl = blockIdx.z;
    const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int m = m0 + ty + 8 * r, l = l0 + tx;
        if (m < p.n_m && l < p.n_l) {
            const int mu = post_bin(p.start_u, l, p.n_u), mv = post_bin(p.start_v, m, p.n_v);
            tile[ty + 8 * r][tx] = ((const CT *)p.spec)[((long long)pl * p.n_v + mv) * p.n_u + mu];
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int l = l0 + ty + 8 * r, m = m0 + tx;
        if (m < p.n_m && l < p.n_l)
            post_pixel<T>(p, tile[tx][ty + 8 * r], l, m, post_bin(p.start_u, l, p.n_u), post_bin(p.start_v, m, p.n_v), p.plane0 + pl);
    }
}


// image -> uv-grid (the inverse of the chain above, for the degridding predict): the cropped, real image is divided by
// the correcting function, zero padded, ifftshift-ed, and multiplied by exp(+2 pi i h j / n) per axis so that the
// forward DFT comes out already fftshift-ed -- one pass writes the FFT input, cuFFT writes the grid.
struct PreParams {
    const void *image;         // real [planes, n_l, n_m]
    void *work;                // complex [planes, n_u, n_v]
    const double2 *phase_u, *phase_v;
    const double *corr_u, *corr_v;
    int n_u, n_v, n_l, n_m, start_u, start_v;
};

template <typename T> __global__ void __launch_bounds__(256) fft_pre_kernel(PreParams p)
{
    using CT = typename Cplx<T>::type;
    const int j1 = blockIdx.x * blockDim.x + threadIdx.x;
    const int j0 = blockIdx.y;
    const int pl = blockIdx.z;
    if (j1 >= p.n_v) return;
    // ifftshift: work[j] = padded[(j + h) % n]
    int i0 = j0 + p.n_u / 2, i1 = j1 + p.n_v / 2;
    if (i0 >= p.n_u) i0 -= p.n_u;
    if (i1 >= p.n_v) i1 -= p.n_v;
    const int l = i0 - p.start_u, m = i1 - p.start_v;
    double val = 0.0;
    if (l >= 0 && l < p.n_l && m >= 0 && m < p.n_m) {
        val = (double)((const T *)p.image)[((long long)pl * p.n_l + l) * p.n_m + m];
        if (p.corr_u) val = val / (p.corr_u[l] * p.corr_v[m]);
    }
    const double2 pu = p.phase_u[j0], pv = p.phase_v[j1];          // conj(pu) conj(pv) = conj(pu pv)
    const double pr = pu.x * pv.x - pu.y * pv.y, pi = -(pu.x * pv.y + pu.y * pv.x);
    CT z;
    z.x = (T)(val * pr), z.y = (T)(val * pi);
    ((CT *)p.work)[((long long)pl * p.n_u + j0) * p.n_v + j1] = z;
}

template <typename T> __global__ void divide_by_centre_kernel(T *image, int n_l, int n_m, long long n_planes, int c_l, int c_m)
{
    // each block handles one plane; the centre value is read before any thread overwrites it
    const long long pl = blockIdx.x;
    if (pl >= n_planes) return;
    T *img = image + pl * (long long)n_l * n_m;
    __shared__ double centre;
    if (threadIdx.x == 0) centre = (double)img[(long long)c_l * n_m + c_m];
    __syncthreads();
    const double c = centre;
    __syncthreads();
    if (c == 0.0 || !isfinite(c)) return;   // e.g. the centre was masked by pb_limit: leave the plane as it is
    for (long long i = threadIdx.x; i < (long long)n_l * n_m; i += blockDim.x) img[i] = (T)((double)img[i] / c);
}

static int make_phase(double2 **dev, int64_t n)
{
    std::vector<double2> h((size_t)n);
    const int64_t half = n / 2;
    for (int64_t m = 0; m < n; ++m) {
        if (n % 2 == 0) {
            h[m] = make_double2((m % 2 == 0) ? 1.0 : -1.0, 0.0);
        } else {
            const int64_t r = (half * m) % n;   // keep the argument small for accuracy
            const double a = -2.0 * M_PI * (double)r / (double)n;
            h[m] = make_double2(cos(a), sin(a));
        }
    }
    CNGI_CUDA_TRY(cudaMalloc((void **)dev, n * sizeof(double2)));
    CNGI_CUDA_TRY(cudaMemcpy(*dev, h.data(), n * sizeof(double2), cudaMemcpyHostToDevice));
    return CNGI_OK;
}

// 2-D transform of nb planes with the shared-memory Bluestein kernel: rows (lines along v) src -> work, then columns
// (lines along u) in place.  dir 0 = unnormalised inverse, 1 = forward.
// keep_v0 / keep_vn: the second pass is only needed for the v bins the caller will read -- the cyclic window
// [keep_v0, keep_v0 + keep_vn) mod n_v (grid_to_image: the cropped image columns); keep_vn <= 0: every bin.
static int blu_transform(cngi_fft_plan *pl, int dir, const float2 *src, float2 *dst, int64_t nb, cudaStream_t st,
                         int keep_v0 = 0, int keep_vn = 0, bool transposed_out = false, const float *src_real = nullptr)
{
    if (!pl->blu_ready[dir]) {
        int rc = blu_axis_create(&pl->blu_u[dir], pl->n_u, dir == 0 ? +1 : -1);
        if (rc == CNGI_OK) rc = blu_axis_create(&pl->blu_v[dir], pl->n_v, dir == 0 ? +1 : -1);
        if (rc != CNGI_OK) return rc;
        pl->blu_ready[dir] = true;
    }
    const long long plane = (long long)pl->n_u * pl->n_v;
    if (keep_vn <= 0 || keep_vn >= pl->n_v) keep_v0 = 0, keep_vn = (int)pl->n_v;
    if (transposed_out) {
        // rows (lines along v) written TRANSPOSED -- strided stores, which nobody waits for, instead of the strided loads a
        // column pass starts with -- then the lines along u are contiguous rows of dst: dst[plane][v bin][u bin]
        // (src must not alias dst here; a real grid is read directly, without a widening pass)
        int rc = blu_lines(pl->blu_v[dir], src, dst, pl->n_v, 1, plane, 1, pl->n_u, plane, (int)pl->n_u, (int)nb, st, 0, 0, src_real);
        if (rc != CNGI_OK) return rc;
        return blu_lines(pl->blu_u[dir], dst, dst, pl->n_u, 1, plane, pl->n_u, 1, plane, keep_vn, (int)nb, st, keep_v0, (int)pl->n_v);
    }
    int rc = blu_lines(pl->blu_v[dir], src, dst, pl->n_v, 1, plane, pl->n_v, 1, plane, (int)pl->n_u, (int)nb, st);
    if (rc != CNGI_OK) return rc;
    return blu_lines(pl->blu_u[dir], dst, dst, 1, pl->n_v, plane, 1, pl->n_v, plane, keep_vn, (int)nb, st, keep_v0, (int)pl->n_v);
}

static int make_cufft(cufftHandle *h, int64_t n_u, int64_t n_v, int64_t batch, int32_t precision)
{
    int n[2] = {(int)n_u, (int)n_v};
    cufftResult r = cufftPlanMany(h, 2, n, nullptr, 1, (int)(n_u * n_v), nullptr, 1, (int)(n_u * n_v),
                                  precision == CNGI_F32 ? CUFFT_C2C : CUFFT_Z2Z, (int)batch);
    if (r != CUFFT_SUCCESS) {
        set_error("cufftPlanMany(%lld x %lld, batch %lld) failed with %d", (long long)n_u, (long long)n_v, (long long)batch, (int)r);
        return CNGI_ERR_CUDA;
    }
    return CNGI_OK;
}

}  // namespace cngi

extern "C" int cngi_b200_fft_plan_create(cngi_fft_plan **out, int64_t n_u, int64_t n_v, int64_t max_planes, int32_t precision)
{
    using namespace cngi;
    CNGI_REQUIRE(out != nullptr, "fft_plan_create: null out pointer");
    CNGI_REQUIRE(n_u > 0 && n_v > 0 && max_planes > 0 && n_u * n_v < (1LL << 31), "fft_plan_create: bad sizes");
    CNGI_REQUIRE(precision == CNGI_F32 || precision == CNGI_F64, "fft_plan_create: bad precision");
    cngi_fft_plan *pl = new cngi_fft_plan();
    pl->n_u = n_u, pl->n_v = n_v, pl->max_planes = max_planes, pl->precision = precision;
    pl->plan = 0, pl->plan_tail = 0, pl->tail_planes = 0, pl->work = nullptr, pl->phase_u = pl->phase_v = nullptr;
    pl->blu_ready[0] = pl->blu_ready[1] = false;
    const char *knob = getenv("CNGI_FFT_BLUESTEIN");   // "0": cuFFT for every size (A/B measurements, tests)
    pl->use_blu = precision == CNGI_F32 && blu_supported(n_u) && blu_supported(n_v) && !(knob && knob[0] == '0');
    int rc = pl->use_blu ? CNGI_OK : make_cufft(&pl->plan, n_u, n_v, max_planes, precision);
    if (rc == CNGI_OK) {
        const size_t cb = precision == CNGI_F32 ? 8 : 16;
        cudaError_t e = cudaMalloc(&pl->work, (size_t)max_planes * n_u * n_v * cb);
        if (e != cudaSuccess) {
            set_error("fft_plan_create: cudaMalloc of the work buffer failed: %s", cudaGetErrorString(e));
            rc = CNGI_ERR_CUDA;
        }
    }
    if (rc == CNGI_OK) rc = make_phase(&pl->phase_u, n_u);
    if (rc == CNGI_OK) rc = make_phase(&pl->phase_v, n_v);
    if (rc != CNGI_OK) {
        cngi_b200_fft_plan_destroy(pl);
        return rc;
    }
    *out = pl;
    return CNGI_OK;
}

extern "C" int cngi_b200_fft_plan_destroy(cngi_fft_plan *pl)
{
    if (!pl) return CNGI_OK;
    if (pl->plan) cufftDestroy(pl->plan);
    if (pl->plan_tail) cufftDestroy(pl->plan_tail);
    for (int d = 0; d < 2; ++d) {
        cngi::blu_axis_destroy(&pl->blu_u[d]);
        cngi::blu_axis_destroy(&pl->blu_v[d]);
    }
    if (pl->work) cudaFree(pl->work);
    if (pl->phase_u) cudaFree(pl->phase_u);
    if (pl->phase_v) cudaFree(pl->phase_v);
    delete pl;
    return CNGI_OK;
}

extern "C" int cngi_b200_grid_to_image(cngi_fft_plan *pl, const cngi_grid_to_image_args *a, void *stream)
{
    using namespace cngi;
    CNGI_REQUIRE(pl && a, "grid_to_image: null plan or args");
    CNGI_REQUIRE(a->grid && a->image, "grid_to_image: null grid or image");
    CNGI_REQUIRE(a->n_u == pl->n_u && a->n_v == pl->n_v && a->precision == pl->precision,
                 "grid_to_image: plan was made for %lld x %lld precision %d", (long long)pl->n_u, (long long)pl->n_v, pl->precision);
    CNGI_REQUIRE(a->image_size[0] > 0 && a->image_size[1] > 0 && a->image_size[0] <= a->n_u && a->image_size[1] <= a->n_v,
                 "grid_to_image: image_size must be within the padded grid");
    CNGI_REQUIRE((a->corr_u == nullptr) == (a->corr_v == nullptr), "grid_to_image: corr_u and corr_v go together");
    CNGI_REQUIRE(a->image_size[0] < 65536, "grid_to_image: image too tall for one launch");
    cudaStream_t st = (cudaStream_t)stream;
    const bool f32 = pl->precision == CNGI_F32;
    const size_t cb = f32 ? 8 : 16, rb = f32 ? 4 : 8;
    const long long plane_cells = (long long)a->n_u * a->n_v;

    for (int64_t p0 = 0; p0 < a->n_planes; p0 += pl->max_planes) {
        const int64_t nb = std::min<int64_t>(pl->max_planes, a->n_planes - p0);
        if (pl->use_blu) {   // shared-memory Bluestein passes (complex64, sides n1 * prime)
            const float2 *bsrc = nullptr;
            const float *rsrc = nullptr;
            if (a->grid_is_complex)
                bsrc = (const float2 *)((const char *)a->grid + (size_t)p0 * plane_cells * cb);
            else
                rsrc = (const float *)a->grid + (size_t)p0 * plane_cells;   // psf grids: the first pass reads the real cells
            // image column m is DFT bin (start_v + m - n_v / 2) mod n_v: the other bins are never read by the post pass
            const int start_v = (int)(a->n_v / 2 - a->image_size[1] / 2);
            const int v0 = (int)(((start_v - a->n_v / 2) % a->n_v + a->n_v) % a->n_v);
            int rc = blu_transform(pl, 0, bsrc, (float2 *)pl->work, nb, st, v0, (int)a->image_size[1], true, rsrc);
            if (rc != CNGI_OK) return rc;
        } else {
            cufftHandle h = pl->plan;
            if (nb != pl->max_planes) {
                if (pl->tail_planes != nb) {
                    if (pl->plan_tail) cufftDestroy(pl->plan_tail);
                    pl->plan_tail = 0, pl->tail_planes = 0;
                    int rc = make_cufft(&pl->plan_tail, pl->n_u, pl->n_v, nb, pl->precision);
                    if (rc != CNGI_OK) return rc;
                    pl->tail_planes = nb;
                }
                h = pl->plan_tail;
            }
            if (cufftSetStream(h, st) != CUFFT_SUCCESS) {
                set_error("cufftSetStream failed");
                return CNGI_ERR_CUDA;
            }
            cufftResult r;
            if (a->grid_is_complex) {
                const char *src = (const char *)a->grid + (size_t)p0 * plane_cells * cb;
                r = f32 ? cufftExecC2C(h, (cufftComplex *)src, (cufftComplex *)pl->work, CUFFT_INVERSE)
                        : cufftExecZ2Z(h, (cufftDoubleComplex *)src, (cufftDoubleComplex *)pl->work, CUFFT_INVERSE);
            } else {
                const char *src = (const char *)a->grid + (size_t)p0 * plane_cells * rb;
                const long long n = nb * plane_cells;
                const unsigned blocks = (unsigned)std::min<long long>(ceil_div(n, 256), (long long)sm_count() * 16);
                if (f32)
                    real_to_complex_kernel<float><<<blocks, 256, 0, st>>>((const float *)src, (float2 *)pl->work, n);
                else
                    real_to_complex_kernel<double><<<blocks, 256, 0, st>>>((const double *)src, (double2 *)pl->work, n);
                CNGI_CUDA_TRY(cudaGetLastError());
                r = f32 ? cufftExecC2C(h, (cufftComplex *)pl->work, (cufftComplex *)pl->work, CUFFT_INVERSE)
                        : cufftExecZ2Z(h, (cufftDoubleComplex *)pl->work, (cufftDoubleComplex *)pl->work, CUFFT_INVERSE);
            }
            if (r != CUFFT_SUCCESS) {
                set_error("cufftExec failed with %d", (int)r);
                return CNGI_ERR_CUDA;
            }
        }
        PostParams pp{};
        pp.spec = pl->work, pp.image = a->image, pp.phase_u = pl->phase_u, pp.phase_v = pl->phase_v;
        pp.sum_weight = a->sum_weight, pp.corr_u = a->corr_u, pp.corr_v = a->corr_v;
        pp.norm_image = a->norm_image, pp.pb_image = a->pb_image;
        pp.norm_planes = a->norm_image_planes, pp.pb_planes = a->pb_image_planes, pp.pb_limit = a->pb_limit;
        pp.n_u = (int)a->n_u, pp.n_v = (int)a->n_v, pp.n_l = (int)a->image_size[0], pp.n_m = (int)a->image_size[1];
        pp.start_u = (int)(a->n_u / 2 - a->image_size[0] / 2), pp.start_v = (int)(a->n_v / 2 - a->image_size[1] / 2);
        pp.plane0 = (int)p0, pp.roundtrip = a->single_precision_roundtrip;
        dim3 grid((unsigned)ceil_div(pp.n_m, 256), (unsigned)pp.n_l, (unsigned)nb);
        CNGI_REQUIRE(nb < 65536, "grid_to_image: too many planes per batch");
        if (pl->use_blu)   // the Bluestein passes leave the spectrum transposed
            fft_post_transposed_kernel<float><<<dim3((unsigned)ceil_div(pp.n_m, 32), (unsigned)ceil_div(pp.n_l, 32), (unsigned)nb),
                                                dim3(32, 8), 0, st>>>(pp);
        else if (f32)
            fft_post_kernel<float><<<grid, 256, 0, st>>>(pp);
        else
            fft_post_kernel<double><<<grid, 256, 0, st>>>(pp);
        CNGI_CUDA_TRY(cudaGetLastError());
    }
    if (a->divide_by_centre) {
        const int c_l = a->divide_by_centre == 2 ? (int)a->centre_pixel[0] : (int)(a->image_size[0] / 2);
        const int c_m = a->divide_by_centre == 2 ? (int)a->centre_pixel[1] : (int)(a->image_size[1] / 2);
        CNGI_REQUIRE(c_l >= 0 && c_l < a->image_size[0] && c_m >= 0 && c_m < a->image_size[1],
                     "grid_to_image: centre pixel (%d, %d) is outside the image", c_l, c_m);
        if (f32)
            divide_by_centre_kernel<float><<<(unsigned)a->n_planes, 256, 0, st>>>((float *)a->image, (int)a->image_size[0],
                                                                                 (int)a->image_size[1], a->n_planes, c_l, c_m);
        else
            divide_by_centre_kernel<double><<<(unsigned)a->n_planes, 256, 0, st>>>((double *)a->image, (int)a->image_size[0],
                                                                                  (int)a->image_size[1], a->n_planes, c_l, c_m);
        CNGI_CUDA_TRY(cudaGetLastError());
    }
    return CNGI_OK;
}

extern "C" int cngi_b200_image_to_grid(cngi_fft_plan *pl, const cngi_image_to_grid_args *a, void *stream)
{
    using namespace cngi;
    CNGI_REQUIRE(pl && a, "image_to_grid: null plan or args");
    CNGI_REQUIRE(a->image && a->grid, "image_to_grid: null image or grid");
    CNGI_REQUIRE(a->n_u == pl->n_u && a->n_v == pl->n_v && a->precision == pl->precision,
                 "image_to_grid: plan was made for %lld x %lld precision %d", (long long)pl->n_u, (long long)pl->n_v, pl->precision);
    CNGI_REQUIRE(a->image_size[0] > 0 && a->image_size[1] > 0 && a->image_size[0] <= a->n_u && a->image_size[1] <= a->n_v,
                 "image_to_grid: image_size must be within the padded grid");
    CNGI_REQUIRE((a->corr_u == nullptr) == (a->corr_v == nullptr), "image_to_grid: corr_u and corr_v go together");
    CNGI_REQUIRE(a->n_u < 65536, "image_to_grid: grid too tall for one launch");
    cudaStream_t st = (cudaStream_t)stream;
    const bool f32 = pl->precision == CNGI_F32;
    const size_t cb = f32 ? 8 : 16, rb = f32 ? 4 : 8;
    const long long plane_cells = (long long)a->n_u * a->n_v, plane_pix = (long long)a->image_size[0] * a->image_size[1];
    for (int64_t p0 = 0; p0 < a->n_planes; p0 += pl->max_planes) {
        const int64_t nb = std::min<int64_t>(pl->max_planes, a->n_planes - p0);
        CNGI_REQUIRE(nb < 65536, "image_to_grid: too many planes per batch");
        cufftHandle h = pl->plan;
        if (!pl->use_blu) {
            if (nb != pl->max_planes) {
                if (pl->tail_planes != nb) {
                    if (pl->plan_tail) cufftDestroy(pl->plan_tail);
                    pl->plan_tail = 0, pl->tail_planes = 0;
                    int rc = make_cufft(&pl->plan_tail, pl->n_u, pl->n_v, nb, pl->precision);
                    if (rc != CNGI_OK) return rc;
                    pl->tail_planes = nb;
                }
                h = pl->plan_tail;
            }
            if (cufftSetStream(h, st) != CUFFT_SUCCESS) {
                set_error("cufftSetStream failed");
                return CNGI_ERR_CUDA;
            }
        }
        PreParams pp{};
        pp.image = (const char *)a->image + (size_t)p0 * plane_pix * rb;
        pp.work = pl->work, pp.phase_u = pl->phase_u, pp.phase_v = pl->phase_v, pp.corr_u = a->corr_u, pp.corr_v = a->corr_v;
        pp.n_u = (int)a->n_u, pp.n_v = (int)a->n_v, pp.n_l = (int)a->image_size[0], pp.n_m = (int)a->image_size[1];
        pp.start_u = (int)(a->n_u / 2 - a->image_size[0] / 2), pp.start_v = (int)(a->n_v / 2 - a->image_size[1] / 2);
        dim3 grid((unsigned)ceil_div(pp.n_v, 256), (unsigned)pp.n_u, (unsigned)nb);
        if (f32)
            fft_pre_kernel<float><<<grid, 256, 0, st>>>(pp);
        else
            fft_pre_kernel<double><<<grid, 256, 0, st>>>(pp);
        CNGI_CUDA_TRY(cudaGetLastError());
        char *dst = (char *)a->grid + (size_t)p0 * plane_cells * cb;
        if (pl->use_blu) {
            int rc = blu_transform(pl, 1, (const float2 *)pl->work, (float2 *)dst, nb, st);
            if (rc != CNGI_OK) return rc;
            continue;
        }
        cufftResult r = f32 ? cufftExecC2C(h, (cufftComplex *)pl->work, (cufftComplex *)dst, CUFFT_FORWARD)
                            : cufftExecZ2Z(h, (cufftDoubleComplex *)pl->work, (cufftDoubleComplex *)dst, CUFFT_FORWARD);
        if (r != CUFFT_SUCCESS) {
            set_error("cufftExec failed with %d", (int)r);
            return CNGI_ERR_CUDA;
        }
    }
    return CNGI_OK;
}
