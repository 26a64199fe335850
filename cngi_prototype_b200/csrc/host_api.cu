// host_api.cu -- host-buffer entry point of the C ABI (what a ctypes/cgo-style binding calls with
// numpy-like HOST arrays).  Mirrors the allocate-and-return contract of _standard_grid_numpy_wrap
// (/root/reference/ngcasa/imaging/_imaging_utils/_standard_grid.py:123-177): the caller hands in host
// vis/uvw/weight/freq/cgk and receives grid + sum_weight in host memory.
//
// The sample arrays are streamed in time chunks through two device staging slots: the H2D copy of
// chunk k+1 (copy stream) overlaps the gridding kernel of chunk k (compute stream).  Time is the
// slowest-varying axis of every sample array, so a chunk is one contiguous byte range per array.
// Device memory comes from the stream-ordered pool (cudaMallocAsync) with the release threshold raised,
// so repeated calls reuse their buffers instead of paying cudaMalloc each time.
#include "common.cuh"
#include <algorithm>

namespace cngi {

struct DevBuf {
    void *p = nullptr;
    cudaStream_t st = nullptr;
    int alloc(size_t bytes, cudaStream_t s)
    {
        st = s;
        CNGI_CUDA_TRY(cudaMallocAsync(&p, bytes ? bytes : 16, s));
        return CNGI_OK;
    }
    ~DevBuf()
    {
        if (p) cudaFreeAsync(p, st);
    }
};

struct StreamPair {
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t loaded[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
    int init()
    {
        CNGI_CUDA_TRY(cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking));
        CNGI_CUDA_TRY(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CNGI_CUDA_TRY(cudaEventCreateWithFlags(&loaded[i], cudaEventDisableTiming));
            CNGI_CUDA_TRY(cudaEventCreateWithFlags(&consumed[i], cudaEventDisableTiming));
        }
        return CNGI_OK;
    }
    ~StreamPair()
    {
        for (int i = 0; i < 2; ++i) {
            if (loaded[i]) cudaEventDestroy(loaded[i]);
            if (consumed[i]) cudaEventDestroy(consumed[i]);
        }
        if (copy) cudaStreamDestroy(copy);
        if (compute) cudaStreamDestroy(compute);
    }
};

}  // namespace cngi

extern "C" int cngi_b200_standard_grid_host(const cngi_std_grid_args *h, int64_t time_chunk)
{
    using namespace cngi;
    CNGI_REQUIRE(h != nullptr, "standard_grid_host: null args");
    CNGI_REQUIRE(h->precision == CNGI_F32 || h->precision == CNGI_F64, "standard_grid_host: bad precision");
    CNGI_REQUIRE(h->weight && h->uvw && h->freq_chan && h->cgk_1D && h->grid && h->sum_weight,
                 "standard_grid_host: null array pointer");
    CNGI_REQUIRE(h->do_psf || h->vis, "standard_grid_host: vis is null in image mode");
    int rc = cngi_b200_check_device();
    if (rc != CNGI_OK) return rc;
    rc = tune_pool_once();
    if (rc != CNGI_OK) return rc;

    const size_t real_b = h->precision == CNGI_F32 ? 4 : 8;
    const size_t cell_b = real_b * (h->complex_grid ? 2 : 1);
    const size_t row = (size_t)h->n_baseline * h->n_chan * h->n_pol;   // samples per time step
    const size_t grid_bytes = (size_t)h->n_imag_chan * h->n_imag_pol * h->n_u * h->n_v * cell_b;
    const size_t sw_bytes = (size_t)h->n_imag_chan * h->n_imag_pol * sizeof(double);
    const int table_len = std::max(1, h->oversampling * (h->support / 2 + 1));
    const bool image = !h->do_psf;

    if (time_chunk <= 0) {   // ~96 MB of visibilities per chunk
        const size_t per_step = row * (image ? 3 * real_b : real_b) + 1;
        time_chunk = (int64_t)std::max<size_t>(1, (96u << 20) / per_step);
    }
    time_chunk = std::min<int64_t>(time_chunk, std::max<int64_t>(h->n_time, 1));

    StreamPair sp;
    rc = sp.init();
    if (rc != CNGI_OK) return rc;

    DevBuf d_grid, d_sw, d_freq, d_cmap, d_pmap, d_cgk, d_vis[2], d_w[2], d_flag[2], d_uvw[2];
    if ((rc = d_grid.alloc(grid_bytes, sp.compute)) || (rc = d_sw.alloc(sw_bytes, sp.compute)) ||
        (rc = d_freq.alloc(h->n_chan * sizeof(double), sp.compute)) ||
        (rc = d_cgk.alloc(table_len * sizeof(double), sp.compute)))
        return rc;
    if (h->chan_map && (rc = d_cmap.alloc(h->n_chan * sizeof(int64_t), sp.compute))) return rc;
    if (h->pol_map && (rc = d_pmap.alloc(h->n_pol * sizeof(int64_t), sp.compute))) return rc;
    for (int s = 0; s < 2; ++s) {
        if (image && (rc = d_vis[s].alloc(time_chunk * row * 2 * real_b, sp.compute))) return rc;
        if ((rc = d_w[s].alloc(time_chunk * row * real_b, sp.compute))) return rc;
        if (image && h->flag && (rc = d_flag[s].alloc(time_chunk * row, sp.compute))) return rc;
        if ((rc = d_uvw[s].alloc(time_chunk * h->n_baseline * 3 * sizeof(double), sp.compute))) return rc;
    }
    CNGI_CUDA_TRY(cudaMemsetAsync(d_grid.p, 0, grid_bytes, sp.compute));
    CNGI_CUDA_TRY(cudaMemsetAsync(d_sw.p, 0, sw_bytes, sp.compute));
    CNGI_CUDA_TRY(cudaMemcpyAsync(d_freq.p, h->freq_chan, h->n_chan * sizeof(double), cudaMemcpyHostToDevice, sp.compute));
    CNGI_CUDA_TRY(cudaMemcpyAsync(d_cgk.p, h->cgk_1D, table_len * sizeof(double), cudaMemcpyHostToDevice, sp.compute));
    if (h->chan_map)
        CNGI_CUDA_TRY(cudaMemcpyAsync(d_cmap.p, h->chan_map, h->n_chan * sizeof(int64_t), cudaMemcpyHostToDevice, sp.compute));
    if (h->pol_map)
        CNGI_CUDA_TRY(cudaMemcpyAsync(d_pmap.p, h->pol_map, h->n_pol * sizeof(int64_t), cudaMemcpyHostToDevice, sp.compute));
    // the staging slots are allocated on the compute stream; the copy stream must not touch them earlier
    CNGI_CUDA_TRY(cudaEventRecord(sp.consumed[0], sp.compute));
    CNGI_CUDA_TRY(cudaEventRecord(sp.consumed[1], sp.compute));

    int chunk_idx = 0;
    for (int64_t t0 = 0; t0 < h->n_time; t0 += time_chunk, ++chunk_idx) {
        const int s = chunk_idx & 1;
        const int64_t nt = std::min<int64_t>(time_chunk, h->n_time - t0);
        const size_t off = (size_t)t0 * row;
        CNGI_CUDA_TRY(cudaStreamWaitEvent(sp.copy, sp.consumed[s], 0));
        if (image)
            CNGI_CUDA_TRY(cudaMemcpyAsync(d_vis[s].p, (const char *)h->vis + off * 2 * real_b, nt * row * 2 * real_b,
                                          cudaMemcpyHostToDevice, sp.copy));
        CNGI_CUDA_TRY(cudaMemcpyAsync(d_w[s].p, (const char *)h->weight + off * real_b, nt * row * real_b,
                                      cudaMemcpyHostToDevice, sp.copy));
        if (image && h->flag)
            CNGI_CUDA_TRY(cudaMemcpyAsync(d_flag[s].p, h->flag + off, nt * row, cudaMemcpyHostToDevice, sp.copy));
        CNGI_CUDA_TRY(cudaMemcpyAsync(d_uvw[s].p, h->uvw + (size_t)t0 * h->n_baseline * 3,
                                      nt * h->n_baseline * 3 * sizeof(double), cudaMemcpyHostToDevice, sp.copy));
        CNGI_CUDA_TRY(cudaEventRecord(sp.loaded[s], sp.copy));

        cngi_std_grid_args d = *h;
        d.n_time = nt;
        d.vis = image ? d_vis[s].p : nullptr;
        d.weight = d_w[s].p;
        d.flag = (image && h->flag) ? (const uint8_t *)d_flag[s].p : nullptr;
        d.uvw = (const double *)d_uvw[s].p;
        d.freq_chan = (const double *)d_freq.p;
        d.chan_map = h->chan_map ? (const int64_t *)d_cmap.p : nullptr;
        d.pol_map = h->pol_map ? (const int64_t *)d_pmap.p : nullptr;
        d.cgk_1D = (const double *)d_cgk.p;
        d.grid = d_grid.p;
        d.sum_weight = (double *)d_sw.p;
        CNGI_CUDA_TRY(cudaStreamWaitEvent(sp.compute, sp.loaded[s], 0));
        rc = cngi_b200_standard_grid(&d, sp.compute);
        if (rc != CNGI_OK) {
            cudaStreamSynchronize(sp.copy);
            cudaStreamSynchronize(sp.compute);
            return rc;
        }
        CNGI_CUDA_TRY(cudaEventRecord(sp.consumed[s], sp.compute));
    }
    CNGI_CUDA_TRY(cudaMemcpyAsync(h->grid, d_grid.p, grid_bytes, cudaMemcpyDeviceToHost, sp.compute));
    CNGI_CUDA_TRY(cudaMemcpyAsync(h->sum_weight, d_sw.p, sw_bytes, cudaMemcpyDeviceToHost, sp.compute));
    CNGI_CUDA_TRY(cudaStreamSynchronize(sp.copy));
    CNGI_CUDA_TRY(cudaStreamSynchronize(sp.compute));
    return CNGI_OK;
}
