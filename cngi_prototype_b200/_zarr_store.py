"""Zarr v2 directory-store arrays without the zarr / numcodecs packages (neither exists in this image).

The reference stores visibilities as `<name>.vis.zarr/<partition>/<VARIABLE>/` directories written by xarray's
to_zarr with `numcodecs.Blosc(cname='zstd', clevel=2, shuffle=0)` (cngi/dio/append_xds.py:69,
cngi/conversion/convert_ms.py) and reads them back with xarray.open_zarr (cngi/dio/read_vis.py:186).  The third-party
pieces restated here are the published formats, not code from /root/reference:

  * zarr storage spec v2 (zarr >= 2.3.2, requirements.txt): `.zarray` JSON {shape, chunks, dtype, compressor, filters,
    fill_value, order, dimension_separator}, one file per chunk named by its chunk indices joined with '.', every
    chunk stored at FULL chunk shape (edge chunks are padded), missing chunk files mean fill_value; xarray adds
    `_ARRAY_DIMENSIONS` to `.zattrs`.
  * c-blosc 1.x frame (numcodecs >= 0.6.3): 16-byte header (version, versionlz, flags, typesize, nbytes, blocksize,
    cbytes), int32 block offsets, blocks of [int32 cbytes, payload] splits; flags 0x1 byte shuffle, 0x2 memcpy,
    0x4 bit shuffle, 0x10 do-not-split, top three bits the codec (1 lz4, 3 zlib, 4 zstd).

Codecs come from the standard library (zlib) and pyarrow (zstd, raw lz4 blocks).  PARITY UNPINNED for the blosc
frames: no zarr/blosc writer exists here and the reference ships no .zarr fixture, so the decoder is checked against
this module's own encoder (round trip) and hand-assembled frames only.  Unsupported encodings (bit shuffle, blosclz,
snappy, filters, object dtypes, Fortran order) raise -- nothing is guessed.
"""
import json
import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_BLOSC_CODECS = {1: "lz4_raw", 3: "zlib", 4: "zstd"}
_BLOSC_IDS = {"lz4": 1, "lz4hc": 1, "zlib": 3, "zstd": 4}
_MAX_SPLITS, _MIN_BUFFERSIZE = 16, 128


class ZarrFormatError(ValueError):
    pass


def _pa_codec(name):
    import pyarrow as pa
    return pa.Codec(name)


def _decompress(codec, payload, n_out):
    """Decoded payload as a bytes-like object of exactly n_out bytes (a pyarrow buffer is not copied again)."""
    if codec == "zlib":
        out = memoryview(zlib.decompress(payload))
    else:
        out = memoryview(_pa_codec(codec).decompress(payload, decompressed_size=n_out)).cast("B")
    if len(out) != n_out:
        raise ZarrFormatError("%s payload decoded to %d bytes, expected %d" % (codec, len(out), n_out))
    return out


def _compress(codec, payload, level):
    if codec == "zlib":
        return zlib.compress(payload, level)
    import pyarrow as pa
    c = pa.Codec(codec) if codec == "lz4_raw" else pa.Codec(codec, compression_level=level)
    return c.compress(payload).to_pybytes()


def _unshuffle(buf, typesize):
    """Inverse of blosc's byte shuffle on one block: the j-th bytes of all elements are stored together."""
    n = len(buf) // typesize
    a = np.frombuffer(buf, dtype=np.uint8)
    body = a[:n * typesize].reshape(typesize, n).T.reshape(-1)
    return body.tobytes() + a[n * typesize:].tobytes()


def _shuffle(buf, typesize):
    n = len(buf) // typesize
    a = np.frombuffer(buf, dtype=np.uint8)
    body = a[:n * typesize].reshape(n, typesize).T.reshape(-1)
    return body.tobytes() + bytes(a[n * typesize:])


def blosc_decode(frame):
    """Decodes one c-blosc 1.x frame (see the module docstring for the layout) into a bytes-like object.

    Every split is decoded straight into its place in one preallocated buffer (one copy after the codec)."""
    frame = memoryview(frame).cast("B")
    if len(frame) < 16:
        raise ZarrFormatError("blosc frame shorter than its header")
    version, _vlz, flags, typesize, nbytes, blocksize, cbytes = struct.unpack_from("<BBBBIII", frame, 0)
    if version != 2:
        raise ZarrFormatError("blosc format version %d is not 2 (c-blosc 1.x)" % version)
    if cbytes != len(frame):
        raise ZarrFormatError("blosc header says %d bytes, chunk file has %d" % (cbytes, len(frame)))
    if flags & 0x4:
        raise NotImplementedError("blosc bit shuffle is not supported")
    if nbytes == 0:
        return b""
    if flags & 0x2:                                    # stored uncompressed
        return frame[16:16 + nbytes]
    codec_id = flags >> 5
    if codec_id not in _BLOSC_CODECS:
        raise NotImplementedError("blosc codec id %d (blosclz / snappy) is not supported" % codec_id)
    codec = _BLOSC_CODECS[codec_id]
    do_shuffle = bool(flags & 0x1) and typesize > 1
    dont_split = bool(flags & 0x10)
    n_blocks = -(-nbytes // blocksize)
    bstarts = struct.unpack_from("<%di" % n_blocks, frame, 16)
    out = bytearray(nbytes)
    view = memoryview(out)
    for b in range(n_blocks):
        bsize = min(blocksize, nbytes - b * blocksize)
        leftover = bsize != blocksize
        split = (not dont_split) and typesize <= _MAX_SPLITS and blocksize // typesize >= _MIN_BUFFERSIZE and not leftover
        n_splits = typesize if split else 1
        ne = bsize // n_splits
        pos = bstarts[b]
        dst = b * blocksize
        for _ in range(n_splits):
            (cb,) = struct.unpack_from("<i", frame, pos)
            pos += 4
            if cb < 0 or pos + cb > len(frame):
                raise ZarrFormatError("blosc split runs past the end of the frame")
            payload = frame[pos:pos + cb]
            pos += cb
            view[dst:dst + ne] = payload if cb == ne else _decompress(codec, payload, ne)
            dst += ne
        if do_shuffle:
            view[b * blocksize:b * blocksize + bsize] = _unshuffle(view[b * blocksize:b * blocksize + bsize], typesize)
    return out


def blosc_encode(raw, typesize, cname="zstd", clevel=2, shuffle=0, blocksize=0):
    """Writes a frame blosc_decode (and c-blosc) reads: unsplit blocks (flag 0x10), optional byte shuffle."""
    raw = bytes(raw)
    if shuffle not in (0, 1):
        raise NotImplementedError("only noshuffle (0) and byte shuffle (1) are written")
    if cname not in _BLOSC_IDS:
        raise NotImplementedError("blosc codec %r is not supported" % cname)
    codec_id = _BLOSC_IDS[cname]
    codec = _BLOSC_CODECS[codec_id]
    nbytes = len(raw)
    blocksize = int(blocksize) or min(max(nbytes, 1), 1 << 18)
    blocksize -= blocksize % typesize if blocksize > typesize else 0
    flags = (codec_id << 5) | 0x10 | (0x1 if (shuffle and typesize > 1) else 0)
    if clevel == 0 or nbytes < 16:
        head = struct.pack("<BBBBIII", 2, 1, flags | 0x2, typesize, nbytes, blocksize, 16 + nbytes)
        return head + raw
    n_blocks = -(-nbytes // blocksize)
    body, bstarts = [], []
    pos = 16 + 4 * n_blocks
    for b in range(n_blocks):
        block = raw[b * blocksize:(b + 1) * blocksize]
        if flags & 0x1:
            block = _shuffle(block, typesize)
        c = _compress(codec, block, clevel)
        if len(c) >= len(block):
            c = block                                  # stored: cbytes == block size
        bstarts.append(pos)
        body.append(struct.pack("<i", len(c)) + c)
        pos += 4 + len(c)
    head = struct.pack("<BBBBIII", 2, 1, flags, typesize, nbytes, blocksize, pos)
    return head + struct.pack("<%di" % n_blocks, *bstarts) + b"".join(body)


def decode_chunk(payload, compressor, typesize):
    if compressor is None:
        return payload
    cid = compressor.get("id")
    if cid == "blosc":
        return blosc_decode(payload)
    if cid == "zlib":
        return zlib.decompress(payload)
    if cid == "gzip":
        return zlib.decompress(payload, 16 + zlib.MAX_WBITS)
    if cid == "zstd":
        import pyarrow as pa
        # a zstd frame carries its content size; pyarrow needs it passed in
        size = _zstd_content_size(payload)
        return pa.Codec("zstd").decompress(payload, decompressed_size=size).to_pybytes()
    raise NotImplementedError("zarr compressor %r is not supported" % (cid,))


def _zstd_content_size(frame):
    """Frame_Content_Size of a zstd frame (RFC 8878 section 3.1.1.1)."""
    if frame[:4] != b"\x28\xb5\x2f\xfd":
        raise ZarrFormatError("not a zstd frame")
    fhd = frame[4]
    fcs_flag, single, did = fhd >> 6, (fhd >> 5) & 1, fhd & 3
    pos = 5 + (0 if single else 1) + (0, 1, 2, 4)[did]
    width = (1 if single else 0, 2, 4, 8)[fcs_flag]
    if width == 0:
        raise ZarrFormatError("zstd frame without a content size")
    val = int.from_bytes(frame[pos:pos + width], "little")
    return val + 256 if width == 2 else val


def encode_chunk(raw, compressor, typesize):
    if compressor is None:
        return raw
    cid = compressor.get("id")
    if cid == "blosc":
        return blosc_encode(raw, typesize, compressor.get("cname", "zstd"), compressor.get("clevel", 2),
                            compressor.get("shuffle", 0), compressor.get("blocksize", 0))
    if cid == "zlib":
        return zlib.compress(raw, compressor.get("level", 1))
    raise NotImplementedError("zarr compressor %r is not written" % (cid,))


def _fill(fill_value, dtype):
    if fill_value is None:
        return np.zeros((), dtype)
    if isinstance(fill_value, str):
        return np.array({"NaN": np.nan, "Infinity": np.inf, "-Infinity": -np.inf}[fill_value], dtype)
    if isinstance(fill_value, list):                   # complex: [re, im]
        re, im = (float(_fill(x, np.float64)) for x in fill_value)
        return np.array(complex(re, im), dtype)
    return np.array(fill_value, dtype)


class ZarrArray:
    """One array of a zarr v2 directory store, read lazily chunk by chunk.

    a.shape / a.chunks / a.dtype / a.dims / a.attrs; a[...] with ints and step-1 slices; numpy.asarray(a) reads it all;
    a.read(region, out=) decodes the intersecting chunks on a thread pool straight into `out` (e.g. a pinned buffer).
    """

    def __init__(self, path):
        self.path = path
        meta_file = os.path.join(path, ".zarray")
        if not os.path.exists(meta_file):
            raise FileNotFoundError("%s is not a zarr v2 array (no .zarray)" % path)
        with open(meta_file) as f:
            m = json.load(f)
        if m.get("zarr_format") != 2:
            raise ZarrFormatError("%s: zarr_format %r is not 2" % (path, m.get("zarr_format")))
        if m.get("order", "C") != "C":
            raise NotImplementedError("%s: Fortran-order chunks are not supported" % path)
        if m.get("filters"):
            raise NotImplementedError("%s: filters %r are not supported" % (path, m["filters"]))
        self.dtype = np.dtype(m["dtype"])
        if self.dtype.hasobject:
            raise NotImplementedError("%s: object dtype" % path)
        self.shape = tuple(int(x) for x in m["shape"])
        self.chunks = tuple(int(x) for x in m["chunks"])
        self.compressor = m.get("compressor")
        self.fill_value = _fill(m.get("fill_value"), self.dtype)
        self.separator = m.get("dimension_separator", ".")
        self.attrs = {}
        attrs_file = os.path.join(path, ".zattrs")
        if os.path.exists(attrs_file):
            with open(attrs_file) as f:
                self.attrs = json.load(f)
        self.dims = tuple(self.attrs.get("_ARRAY_DIMENSIONS", ()))
        self.ndim = len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    def __len__(self):
        return self.shape[0]

    def chunk_file(self, idx):
        return os.path.join(self.path, self.separator.join(str(i) for i in idx) if idx else "0")

    def read_chunk(self, idx):
        """The full-shape chunk `idx` (edge chunks keep their padding); fill_value when the file is absent."""
        f = self.chunk_file(idx)
        if not os.path.exists(f):
            return np.full(self.chunks, self.fill_value, dtype=self.dtype)
        with open(f, "rb") as fh:
            payload = fh.read()
        raw = decode_chunk(payload, self.compressor, self.dtype.itemsize)
        want = int(np.prod(self.chunks, dtype=np.int64)) * self.dtype.itemsize
        if len(raw) != want:
            raise ZarrFormatError("%s: chunk decodes to %d bytes, expected %d" % (f, len(raw), want))
        return np.frombuffer(raw, dtype=self.dtype).reshape(self.chunks)

    def _normalise(self, region):
        if region is None:
            region = ()
        if not isinstance(region, tuple):
            region = (region,)
        if any(r is Ellipsis for r in region):
            k = region.index(Ellipsis)
            region = region[:k] + (slice(None),) * (self.ndim - len(region) + 1) + region[k + 1:]
        region = region + (slice(None),) * (self.ndim - len(region))
        if len(region) != self.ndim:
            raise IndexError("too many indices for a %d-d zarr array" % self.ndim)
        bounds, squeeze = [], []
        for ax, (r, n) in enumerate(zip(region, self.shape)):
            if isinstance(r, (int, np.integer)):
                i = int(r) + (n if r < 0 else 0)
                if not 0 <= i < n:
                    raise IndexError("index %d is out of bounds for axis %d with size %d" % (r, ax, n))
                bounds.append((i, i + 1))
                squeeze.append(ax)
            elif isinstance(r, slice):
                lo, hi, step = r.indices(n)
                if step != 1:
                    raise NotImplementedError("strided reads are not supported")
                bounds.append((lo, max(lo, hi)))
            else:
                raise NotImplementedError("only ints and slices index a zarr array")
        return bounds, tuple(squeeze)

    def _native_compressor(self):
        from . import _lib
        cid = None if self.compressor is None else self.compressor.get("id")
        return {None: _lib.ZARR_RAW, "zlib": _lib.ZARR_ZLIB, "blosc": _lib.ZARR_BLOSC}.get(cid)

    def _read_native(self, bounds, out, threads):
        """The region through cngi_b200_zarr_read_chunks (csrc/zarr_chunk_reader.cu): file reads, decoding and the box
        copies run on native threads without the GIL, straight into `out` (e.g. a pinned staging buffer)."""
        import ctypes as C
        from . import _lib
        comp = self._native_compressor()
        if comp is None or self.ndim > 8 or not out.flags.c_contiguous:
            raise NotImplementedError("native chunk reader: compressor %r / layout not supported" % (self.compressor,))
        ranges = [range(lo // c, (hi - 1) // c + 1) for (lo, hi), c in zip(bounds, self.chunks)]
        todo = [()]
        for r in ranges:
            todo = [t + (i,) for t in todo for i in r]
        jobs = (_lib.ZarrChunkJob * len(todo))()
        for job, idx in zip(jobs, todo):
            f = self.chunk_file(idx)
            job.path = f.encode() if os.path.exists(f) else None
            for ax, (i, c, (lo, hi)) in enumerate(zip(idx, self.chunks, bounds)):
                a, b = max(lo, i * c), min(hi, (i + 1) * c)
                job.chunk_shape[ax], job.src_start[ax], job.dst_start[ax], job.extent[ax] = c, a - i * c, a - lo, b - a
        shape = (C.c_int64 * max(self.ndim, 1))(*out.shape)
        fill = np.ascontiguousarray(self.fill_value.astype(self.dtype))
        rc = _lib.lib().cngi_b200_zarr_read_chunks(jobs, len(todo), out.ctypes.data, shape, self.ndim, self.dtype.itemsize,
                                                   comp, fill.ctypes.data, int(threads))
        _lib.check(rc, "cngi_b200_zarr_read_chunks")

    def read(self, region=None, out=None, pool=None, threads=0):
        """threads > 0: decode on that many native threads (libcngi_b200.so); otherwise in Python (`pool` optional)."""
        bounds, squeeze = self._normalise(region)
        shape = tuple(hi - lo for lo, hi in bounds)
        if out is None:
            out = np.empty(shape, dtype=self.dtype)
        elif tuple(out.shape) != shape or out.dtype != self.dtype:
            raise ValueError("out has shape %s / dtype %s, the region needs %s / %s"
                             % (tuple(out.shape), out.dtype, shape, self.dtype))
        if out.size and threads:
            self._read_native(bounds, out, threads)
        elif out.size:
            ranges = [range(lo // c, (hi - 1) // c + 1) for (lo, hi), c in zip(bounds, self.chunks)]
            todo = [()]
            for r in ranges:
                todo = [t + (i,) for t in todo for i in r]

            def one(idx):
                chunk = self.read_chunk(idx)
                src, dst = [], []
                for i, c, (lo, hi) in zip(idx, self.chunks, bounds):
                    a, b = max(lo, i * c), min(hi, (i + 1) * c)
                    src.append(slice(a - i * c, b - i * c))
                    dst.append(slice(a - lo, b - lo))
                out[tuple(dst)] = chunk[tuple(src)]

            if pool is not None and len(todo) > 1:
                list(pool.map(one, todo))              # zlib / pyarrow release the GIL while decoding
            else:
                for idx in todo:
                    one(idx)
        return out.reshape([n for ax, n in enumerate(shape) if ax not in squeeze]) if squeeze else out

    def __getitem__(self, region):
        return self.read(region)

    def __array__(self, dtype=None, copy=None):
        a = self.read()
        return a if dtype is None else a.astype(dtype, copy=False)

    def __repr__(self):
        return "ZarrArray(%s, shape=%s, chunks=%s, dtype=%s)" % (self.path, self.shape, self.chunks, self.dtype)


def write_array(path, array, chunks=None, compressor=None, dims=None, attrs=None, fill_value=None):
    """Writes `array` as a zarr v2 array directory (what xarray.to_zarr lays down for one variable)."""
    array = np.asarray(array)
    if array.dtype.hasobject:
        raise NotImplementedError("object dtype")
    chunks = tuple(int(c) for c in (chunks or array.shape))
    chunks = tuple(max(1, min(c, n) if n else 1) for c, n in zip(chunks, array.shape)) if array.ndim else ()
    os.makedirs(path, exist_ok=True)
    dtype_str = array.dtype.str if array.dtype.itemsize > 1 else "|" + array.dtype.str[1:]
    if isinstance(fill_value, float) and np.isnan(fill_value):
        fill_json = "NaN"
    elif isinstance(fill_value, complex):
        fill_json = [("NaN" if np.isnan(x) else x) for x in (fill_value.real, fill_value.imag)]
    else:
        fill_json = fill_value
    meta = {"zarr_format": 2, "shape": list(array.shape), "chunks": list(chunks), "dtype": dtype_str,
            "compressor": compressor, "fill_value": fill_json, "filters": None, "order": "C"}
    with open(os.path.join(path, ".zarray"), "w") as f:
        json.dump(meta, f, indent=4)
    za = dict(attrs or {})
    if dims is not None:
        za["_ARRAY_DIMENSIONS"] = list(dims)
    with open(os.path.join(path, ".zattrs"), "w") as f:
        json.dump(za, f, indent=4)
    grid = [range(-(-n // c)) for n, c in zip(array.shape, chunks)]
    todo = [()]
    for r in grid:
        todo = [t + (i,) for t in todo for i in r]
    for idx in todo:
        block = np.zeros(chunks, dtype=array.dtype)    # edge chunks are stored at full chunk shape
        sl = tuple(slice(i * c, min(n, (i + 1) * c)) for i, c, n in zip(idx, chunks, array.shape))
        piece = array[sl]
        block[tuple(slice(0, s) for s in piece.shape)] = piece
        name = ".".join(str(i) for i in idx) if idx else "0"
        with open(os.path.join(path, name), "wb") as f:
            f.write(encode_chunk(block.tobytes(), compressor, array.dtype.itemsize))
    return path


def open_group(path):
    """{name: ZarrArray} for the arrays directly under a zarr v2 group directory, plus the group's .zattrs."""
    if not os.path.isdir(path):
        raise FileNotFoundError(path)
    arrays = {}
    for name in sorted(os.listdir(path)):
        sub = os.path.join(path, name)
        if os.path.isdir(sub) and os.path.exists(os.path.join(sub, ".zarray")):
            arrays[name] = sub
    attrs = {}
    if os.path.exists(os.path.join(path, ".zattrs")):
        with open(os.path.join(path, ".zattrs")) as f:
            attrs = json.load(f)
    return arrays, attrs


def make_pool(workers=8):
    return ThreadPoolExecutor(max_workers=int(workers), thread_name_prefix="zarr-decode")
