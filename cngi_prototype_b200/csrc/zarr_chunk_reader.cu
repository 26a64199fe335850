// zarr_chunk_reader.cu -- N4, host side: decode zarr v2 chunk files into a region of a (pinned) host array on native
// threads.  No device code; it lives in libcngi_b200.so so that the Python mirror (read_vis.py) can fill the pinned
// staging buffers of the chunk stream without the GIL.
//
// What the reference does here: xarray.open_zarr + dask read one chunk per task through zarr-python / numcodecs
// (cngi/dio/read_vis.py:186-197; chunks written with Blosc(cname='zstd', clevel=2, shuffle=0), cngi/dio/append_xds.py:69).
// Formats restated from their published layouts (zarr storage spec v2; c-blosc 1.x frame header, block offsets and
// [int32 cbytes, payload] splits; LZ4 block format) -- not from /root/reference, which holds none of this code.
// zstd and zlib payloads are decoded by the system's libzstd.so.1 / libz.so.1, opened with dlopen (the image ships the
// runtime libraries without headers; the two prototypes used are part of their stable public ABI).
#include "common.cuh"

#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>

#include <atomic>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace cngi {
namespace {

typedef size_t (*zstd_decompress_fn)(void *, size_t, const void *, size_t);
typedef unsigned (*zstd_is_error_fn)(size_t);
typedef int (*zlib_uncompress_fn)(unsigned char *, unsigned long *, const unsigned char *, unsigned long);

struct Codecs {
    zstd_decompress_fn zstd_decompress = nullptr;
    zstd_is_error_fn zstd_is_error = nullptr;
    zlib_uncompress_fn zlib_uncompress = nullptr;
};

const Codecs &codecs()
{
    static Codecs c;
    static std::once_flag once;
    std::call_once(once, [] {
        if (void *h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_GLOBAL)) {
            c.zstd_decompress = (zstd_decompress_fn)dlsym(h, "ZSTD_decompress");
            c.zstd_is_error = (zstd_is_error_fn)dlsym(h, "ZSTD_isError");
        }
        if (void *h = dlopen("libz.so.1", RTLD_NOW | RTLD_GLOBAL)) c.zlib_uncompress = (zlib_uncompress_fn)dlsym(h, "uncompress");
    });
    return c;
}

// LZ4 block format: sequences of [token][literal length+][literals][offset16][match length+], the last one literals only.
bool lz4_block_decode(const uint8_t *src, size_t n, uint8_t *dst, size_t cap)
{
    const uint8_t *ip = src, *iend = src + n;
    uint8_t *op = dst, *oend = dst + cap;
    while (ip < iend) {
        const unsigned token = *ip++;
        size_t lit = token >> 4;
        if (lit == 15) {
            unsigned b;
            do {
                if (ip >= iend) return false;
                b = *ip++;
                lit += b;
            } while (b == 255);
        }
        if (lit > (size_t)(iend - ip) || lit > (size_t)(oend - op)) return false;
        memcpy(op, ip, lit);
        op += lit;
        ip += lit;
        if (ip >= iend) break;
        if (iend - ip < 2) return false;
        const size_t off = (size_t)ip[0] | ((size_t)ip[1] << 8);
        ip += 2;
        if (off == 0 || off > (size_t)(op - dst)) return false;
        size_t ml = token & 15;
        if (ml == 15) {
            unsigned b;
            do {
                if (ip >= iend) return false;
                b = *ip++;
                ml += b;
            } while (b == 255);
        }
        ml += 4;
        if (ml > (size_t)(oend - op)) return false;
        const uint8_t *m = op - off;
        if (off >= ml) memcpy(op, m, ml);
        else for (size_t k = 0; k < ml; ++k) op[k] = m[k];   // overlapping run
        op += ml;
    }
    return op == oend;
}

// Scratch that grows without being zero-filled (std::vector::resize would add a memset pass over every chunk).
struct Scratch {
    std::unique_ptr<uint8_t[]> p;
    size_t cap = 0, len = 0;
    void resize(size_t n)
    {
        if (n > cap) {
            p.reset(new uint8_t[n]);
            cap = n;
        }
        len = n;
    }
    uint8_t *data() { return p.get(); }
};

enum { CODEC_LZ4 = 1, CODEC_ZLIB = 3, CODEC_ZSTD = 4 };

// payload -> exactly n_out bytes at dst.  Returns an error text or nullptr.
const char *codec_decode(int codec, const uint8_t *payload, size_t n_in, uint8_t *dst, size_t n_out)
{
    const Codecs &c = codecs();
    if (codec == CODEC_ZSTD) {
        if (!c.zstd_decompress) return "libzstd.so.1 is not available";
        const size_t r = c.zstd_decompress(dst, n_out, payload, n_in);
        if (c.zstd_is_error(r) || r != n_out) return "zstd payload does not decode to the expected size";
        return nullptr;
    }
    if (codec == CODEC_ZLIB) {
        if (!c.zlib_uncompress) return "libz.so.1 is not available";
        unsigned long len = (unsigned long)n_out;
        if (c.zlib_uncompress(dst, &len, payload, (unsigned long)n_in) != 0 || len != n_out)
            return "zlib payload does not decode to the expected size";
        return nullptr;
    }
    if (codec == CODEC_LZ4) return lz4_block_decode(payload, n_in, dst, n_out) ? nullptr : "corrupt lz4 block";
    return "blosc codec is not supported (blosclz / snappy)";
}

uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// One c-blosc 1.x frame -> dst (exactly n_out bytes).  tmp is scratch for shuffled blocks.
const char *blosc_decode(const uint8_t *f, size_t n, uint8_t *dst, size_t n_out, Scratch &tmp)
{
    if (n < 16) return "blosc frame shorter than its header";
    const unsigned version = f[0], flags = f[2], typesize = f[3];
    const size_t nbytes = rd32(f + 4), blocksize = rd32(f + 8), cbytes = rd32(f + 12);
    if (version != 2) return "blosc format version is not 2 (c-blosc 1.x)";
    if (cbytes != n) return "blosc header size differs from the chunk file size";
    if (nbytes != n_out) return "blosc frame does not hold one full chunk";
    if (flags & 0x4) return "blosc bit shuffle is not supported";
    if (nbytes == 0) return nullptr;
    if (flags & 0x2) {
        if (16 + nbytes > n) return "truncated blosc memcpy frame";
        memcpy(dst, f + 16, nbytes);
        return nullptr;
    }
    if (blocksize == 0 || typesize == 0) return "corrupt blosc header";
    const int codec = (int)(flags >> 5);
    const bool shuffle = (flags & 0x1) && typesize > 1, dont_split = (flags & 0x10) != 0;
    const size_t n_blocks = (nbytes + blocksize - 1) / blocksize;
    if (16 + 4 * n_blocks > n) return "truncated blosc block table";
    for (size_t b = 0; b < n_blocks; ++b) {
        const size_t bsize = (b + 1) * blocksize <= nbytes ? blocksize : nbytes - b * blocksize;
        const bool leftover = bsize != blocksize;
        const bool split = !dont_split && typesize <= 16 && blocksize / typesize >= 128 && !leftover;
        const size_t n_splits = split ? typesize : 1, ne = bsize / n_splits;
        size_t pos = rd32(f + 16 + 4 * b);
        uint8_t *blk = dst + b * blocksize;
        if (shuffle) {
            tmp.resize(bsize);
            blk = tmp.data();
        }
        for (size_t s = 0; s < n_splits; ++s) {
            if (pos + 4 > n) return "blosc split runs past the end of the frame";
            const int32_t cb = (int32_t)rd32(f + pos);
            pos += 4;
            if (cb < 0 || pos + (size_t)cb > n) return "blosc split runs past the end of the frame";
            if ((size_t)cb == ne) memcpy(blk + s * ne, f + pos, ne);
            else if (const char *e = codec_decode(codec, f + pos, (size_t)cb, blk + s * ne, ne)) return e;
            pos += (size_t)cb;
        }
        if (shuffle) {   // byte j of element i was stored at j * n_elem + i
            uint8_t *out = dst + b * blocksize;
            const size_t n_elem = bsize / typesize;
            for (size_t j = 0; j < typesize; ++j) {
                const uint8_t *src = blk + j * n_elem;
                for (size_t i = 0; i < n_elem; ++i) out[i * typesize + j] = src[i];
            }
            memcpy(out + n_elem * typesize, blk + n_elem * typesize, bsize - n_elem * typesize);
        }
    }
    return nullptr;
}

// A chunk file mapped read-only: stored (uncompressed) payloads are then copied once, page cache -> destination.
struct MappedFile {
    const uint8_t *p = nullptr;
    size_t n = 0;
    bool open(const char *path, bool &missing)
    {
        missing = false;
        const int fd = ::open(path, O_RDONLY);
        if (fd < 0) {
            missing = errno == ENOENT;
            return missing;
        }
        struct stat st;
        bool ok = fstat(fd, &st) == 0;
        if (ok && st.st_size > 0) {
            void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
            ok = m != MAP_FAILED;
            if (ok) {
                p = (const uint8_t *)m;
                n = (size_t)st.st_size;
            }
        }
        ::close(fd);
        return ok;
    }
    void close()
    {
        if (p) munmap((void *)p, n);
        p = nullptr;
        n = 0;
    }
    ~MappedFile() { close(); }
    const uint8_t *data() const { return p; }
    size_t size() const { return n; }
};

struct Reader {
    const cngi_zarr_chunk_job *jobs;
    int64_t n_jobs;
    uint8_t *dst;
    const int64_t *dst_shape;
    int ndim, elem;
    int compressor;
    const uint8_t *fill;
    std::atomic<int64_t> next{0};
    std::mutex mu;
    std::string error;

    void fail(const std::string &msg)
    {
        std::lock_guard<std::mutex> g(mu);
        if (error.empty()) error = msg;
    }

    void run()
    {
        Scratch chunk, tmp;
        for (;;) {
            const int64_t j = next.fetch_add(1);
            if (j >= n_jobs) return;
            {
                std::lock_guard<std::mutex> g(mu);
                if (!error.empty()) return;
            }
            const cngi_zarr_chunk_job &job = jobs[j];
            size_t chunk_elems = 1;
            for (int d = 0; d < ndim; ++d) chunk_elems *= (size_t)job.chunk_shape[d];
            const size_t chunk_bytes = chunk_elems * elem;
            bool missing = true;
            MappedFile file;
            if (job.path && !file.open(job.path, missing)) return fail(std::string("cannot read ") + job.path);
            const uint8_t *src = nullptr;
            if (!missing) {
                const char *e = nullptr;
                if (compressor == CNGI_ZARR_RAW) {
                    if (file.size() != chunk_bytes) e = "raw chunk file has the wrong size";
                    src = file.data();
                } else if (compressor == CNGI_ZARR_BLOSC && file.size() == 16 + chunk_bytes && file.size() >= 16 &&
                           (file.data()[2] & 0x2) && file.data()[0] == 2 && rd32(file.data() + 4) == chunk_bytes &&
                           rd32(file.data() + 12) == file.size()) {
                    src = file.data() + 16;   // blosc kept the chunk uncompressed (memcpy frame): copy it once
                } else {
                    chunk.resize(chunk_bytes);
                    src = chunk.data();
                    if (compressor == CNGI_ZARR_BLOSC) e = blosc_decode(file.data(), file.size(), chunk.data(), chunk_bytes, tmp);
                    else if (compressor == CNGI_ZARR_ZLIB) e = codec_decode(CODEC_ZLIB, file.data(), file.size(), chunk.data(), chunk_bytes);
                    else e = "unknown compressor id";
                }
                if (e) return fail(std::string(job.path) + ": " + e);
            }
            copy_box(job, src);
        }
    }

    // Copies job.extent from the decoded chunk (or the fill element) into the destination region, innermost axis as
    // one memcpy per row.
    void copy_box(const cngi_zarr_chunk_job &job, const uint8_t *src)
    {
        int64_t sstride[8], dstride[8];
        int64_t ss = elem, ds = elem;
        for (int d = ndim - 1; d >= 0; --d) {
            sstride[d] = ss;
            dstride[d] = ds;
            ss *= job.chunk_shape[d];
            ds *= dst_shape[d];
        }
        int64_t rows = 1;
        for (int d = 0; d < ndim - 1; ++d) rows *= job.extent[d];
        const int64_t run = ndim ? job.extent[ndim - 1] : 1;
        if (run <= 0 || rows <= 0) return;
        int64_t idx[8] = {0};
        for (int64_t r = 0; r < rows; ++r) {
            int64_t so = ndim ? (job.src_start[ndim - 1]) * sstride[ndim - 1] : 0;
            int64_t dof = ndim ? (job.dst_start[ndim - 1]) * dstride[ndim - 1] : 0;
            for (int d = 0; d < ndim - 1; ++d) {
                so += (job.src_start[d] + idx[d]) * sstride[d];
                dof += (job.dst_start[d] + idx[d]) * dstride[d];
            }
            if (src) memcpy(dst + dof, src + so, (size_t)run * elem);
            else for (int64_t k = 0; k < run; ++k) memcpy(dst + dof + k * elem, fill, elem);
            for (int d = ndim - 2; d >= 0; --d) {
                if (++idx[d] < job.extent[d]) break;
                idx[d] = 0;
            }
        }
    }
};

}  // namespace
}  // namespace cngi

extern "C" int cngi_b200_zarr_read_chunks(const cngi_zarr_chunk_job *jobs, int64_t n_jobs, void *dst_host,
                                          const int64_t *dst_shape, int32_t ndim, int32_t elem_bytes, int32_t compressor,
                                          const void *fill_elem, int32_t n_threads)
{
    using namespace cngi;
    CNGI_REQUIRE(n_jobs >= 0 && ndim >= 0 && ndim <= 8, "zarr_read_chunks: bad n_jobs / ndim");
    if (n_jobs == 0) return CNGI_OK;
    CNGI_REQUIRE(jobs && dst_host && (ndim == 0 || dst_shape) && fill_elem, "zarr_read_chunks: NULL argument");
    CNGI_REQUIRE(elem_bytes >= 1 && elem_bytes <= 64, "zarr_read_chunks: bad element size");
    CNGI_REQUIRE(compressor == CNGI_ZARR_RAW || compressor == CNGI_ZARR_ZLIB || compressor == CNGI_ZARR_BLOSC,
                 "zarr_read_chunks: compressor is not one of CNGI_ZARR_RAW / ZLIB / BLOSC");
    for (int64_t j = 0; j < n_jobs; ++j)
        for (int d = 0; d < ndim; ++d) {
            const cngi_zarr_chunk_job &q = jobs[j];
            CNGI_REQUIRE(q.chunk_shape[d] >= 1 && q.extent[d] >= 0 && q.src_start[d] >= 0 && q.dst_start[d] >= 0 &&
                             q.src_start[d] + q.extent[d] <= q.chunk_shape[d] && q.dst_start[d] + q.extent[d] <= dst_shape[d],
                         "zarr_read_chunks: job %lld leaves its chunk or the destination on axis %d", (long long)j, d);
        }
    Reader R;
    R.jobs = jobs;
    R.n_jobs = n_jobs;
    R.dst = (uint8_t *)dst_host;
    R.dst_shape = dst_shape;
    R.ndim = ndim;
    R.elem = elem_bytes;
    R.compressor = compressor;
    R.fill = (const uint8_t *)fill_elem;
    int nt = n_threads < 1 ? 1 : (n_threads > 64 ? 64 : n_threads);
    if ((int64_t)nt > n_jobs) nt = (int)n_jobs;
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back([&R] { R.run(); });
    R.run();
    for (auto &t : pool) t.join();
    if (!R.error.empty()) {
        set_error("zarr_read_chunks: %s", R.error.c_str());
        return CNGI_ERR_INVALID;
    }
    return CNGI_OK;
}
