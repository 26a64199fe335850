"""read_vis: zarr visibility datasets -> pinned host buffers -> device chunks.

Mirrors cngi/dio/read_vis.py:21 (local-filesystem branch :182-197): `read_vis(infile, partition=None, chunks=None)`
opens every partition directory of a `.vis.zarr` store ('global/*' sub-tables included when 'global' is asked for) and
returns a dataset of datasets -- `mxds.attrs[partition]` is one visibility dataset, as vis_xds_packager builds it
(cngi/_utils/_io.py:37).  xarray / dask do not exist in this image, so a dataset here is `VisDataset`: a read-only
mapping {variable or coordinate name -> lazy ZarrArray} with `.dims`, `.attrs` and `.chunks`; `numpy.asarray(xds[k])`
reads a variable, and the imaging API (imaging.make_image etc.) accepts it as is.  The S3 branch (:63-180) is a
network path and is not built.

What the reference leaves to dask (one task per zarr chunk, host arrays shipped between workers) is here a
software pipeline per GPU: `VisDataset.iter_device_chunks()` decodes the zarr chunks of the next time block on a
thread pool straight into PINNED host buffers while the previous block is copied host -> device on a copy stream and
the one before that is gridded; every yielded block is a dict of device tensors ordered behind its copy by a CUDA
event.  `write_vis` lays down the same directory structure (for tests, fixtures and synthetic benchmarks).
"""
import json
import os
import threading
from collections.abc import Mapping

import numpy as np

from . import _zarr_store as zs

DEFAULT_COMPRESSOR = {"id": "blosc", "cname": "zstd", "clevel": 2, "shuffle": 0, "blocksize": 0}   # append_xds.py:69
SAMPLE_DIMS = ("time", "baseline", "chan", "pol")


class VisDataset(Mapping):
    """One partition of a vis.zarr store: lazy variables and coordinates, chunked as they are on disk."""

    def __init__(self, path, chunks=None):
        self.path = path
        paths, self.attrs = zs.open_group(path)
        self._arrays = {}
        for k, p in paths.items():
            try:
                self._arrays[k] = zs.ZarrArray(p)
            except NotImplementedError:      # string tables (object dtype + vlen filter): not needed for imaging
                pass
        self.dims = {}
        for a in self._arrays.values():
            for d, n in zip(a.dims, a.shape):
                self.dims[d] = n
        self._chunks_override = dict(chunks or {})

    # ---- mapping protocol ----------------------------------------------------------------------------------
    def __getitem__(self, key):
        if key == "chunks":
            return self.chunks
        return self._arrays[key]

    def __iter__(self):
        return iter(self._arrays)

    def __len__(self):
        return len(self._arrays)

    def get(self, key, default=None):
        if key == "chunks":
            return self.chunks
        return self._arrays.get(key, default)

    @property
    def data_vars(self):
        return [k for k, a in self._arrays.items() if k not in self.dims]

    @property
    def chunks(self):
        """{'time': n, 'baseline': n, 'chan': n, 'pol': n}: the zarr chunking of DATA (or of the first 4-d variable),
        overridden by read_vis(chunks=) -- what `IMAGING_WEIGHT.data.numblocks` encodes in the reference
        (_standard_grid.py:35)."""
        sample_vars = [a for a in self._arrays.values() if a.dims == SAMPLE_DIMS]
        first = self._arrays["DATA"] if "DATA" in self._arrays else (sample_vars[0] if sample_vars else None)
        out = dict(zip(first.dims, first.chunks)) if first is not None else {}
        for a in self._arrays.values():
            for d, c in zip(a.dims, a.chunks):
                out.setdefault(d, c)
        out.update(self._chunks_override)
        return out

    def load(self, names=None):
        """{name: numpy array} of the named (default: all) variables, fully read."""
        return {k: np.asarray(self._arrays[k]) for k in (names or list(self._arrays))}

    # ---- host-side chunk walk --------------------------------------------------------------------------------
    def time_blocks(self, time_chunk=0):
        n_time = int(self.dims["time"])
        step = int(time_chunk) or int(self.chunks.get("time", n_time)) or n_time
        return [slice(t, min(n_time, t + step)) for t in range(0, n_time, max(step, 1))]

    def _names(self, names):
        names = list(names) if names is not None else [k for k in ("DATA", "UVW", "WEIGHT", "DATA_WEIGHT", "FLAG",
                                                                   "IMAGING_WEIGHT", "FIELD_ID") if k in self._arrays]
        for k in names:
            a = self._arrays[k]
            if not a.dims or a.dims[0] != "time":
                raise ValueError("%s has dims %s: only variables with a leading time axis are streamed" % (k, a.dims))
        return names

    def iter_host_chunks(self, names=None, time_chunk=0, workers=8, native=True):
        """Yields (time slice, {name: numpy array}) per time block; chunk files are decoded on `workers` threads --
        native ones (cngi_b200_zarr_read_chunks, no GIL) or, with native=False, a Python thread pool."""
        names = self._names(names)
        with zs.make_pool(workers) as pool:
            for sl in self.time_blocks(time_chunk):
                if native:
                    yield sl, {k: self._arrays[k].read((sl,), threads=workers) for k in names}
                else:
                    yield sl, {k: self._arrays[k].read((sl,), pool=pool) for k in names}

    # ---- device pipeline -------------------------------------------------------------------------------------
    def iter_device_chunks(self, names=None, time_chunk=0, device=None, workers=8, depth=2, native=True):
        """Yields (time slice, {name: CUDA tensor}) per time block.

        Three stages overlap: zarr decode into pinned buffers (reader thread + `workers` native decode threads,
        cngi_b200_zarr_read_chunks; native=False decodes in Python instead), H2D on a
        copy stream, and the caller's kernels on the current stream.  `depth` buffer sets (pinned + device) rotate:
        the reader refills a pinned set once its copy has landed, and the copy stream overwrites a device set only
        behind the event recorded after the caller queued its kernels on it (i.e. when the caller asks for the next
        block -- use a yielded block before advancing the iterator, or clone it).  bool variables arrive as uint8
        (what the gridders' flag argument takes).
        """
        from ._devutil import torch
        from . import _lib
        _lib.require_device()
        names = self._names(names)
        blocks = self.time_blocks(time_chunk)
        if not blocks:
            return
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        depth = max(2, int(depth))
        max_t = max(sl.stop - sl.start for sl in blocks)

        def np_dtype(a):
            return np.dtype(np.uint8) if a.dtype == np.bool_ else a.dtype

        pinned = [{k: torch.empty((max_t,) + self._arrays[k].shape[1:], dtype=torch.from_numpy(
            np.empty(0, np_dtype(self._arrays[k]))).dtype, pin_memory=True) for k in names} for _ in range(depth)]
        on_dev = [{k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in pinned[0].items()} for _ in range(depth)]
        copy_stream = torch.cuda.Stream(device=dev)
        copied = [None] * depth          # event: H2D of the set finished (recorded on the copy stream)
        consumed = [None] * depth        # event: the caller's kernels on the set were queued before this point
        host_free = [threading.Semaphore(1) for _ in range(depth)]
        host_full = [threading.Semaphore(0) for _ in range(depth)]
        failure = []

        stop = threading.Event()

        def reader():
            try:
                with zs.make_pool(workers) as pool, zs.make_pool(len(names)) as var_pool:
                    for i, sl in enumerate(blocks):
                        s = i % depth
                        host_free[s].acquire()
                        if stop.is_set():
                            return
                        n = sl.stop - sl.start

                        def fill(k, s=s, n=n, sl=sl):
                            a = self._arrays[k]
                            out = pinned[s][k].numpy()[:n]
                            out = out.view(np.bool_) if a.dtype == np.bool_ else out
                            if native:
                                a.read((sl,), out=out, threads=workers)
                            else:
                                a.read((sl,), out=out, pool=pool)

                        if native:                     # the variables of a block are read concurrently (no GIL inside)
                            list(var_pool.map(fill, names))
                        else:
                            for k in names:
                                fill(k)
                        host_full[s].release()
            except BaseException as e:          # surfaced in the consumer; never swallowed
                failure.append(e)
                for sem in host_full:
                    sem.release()

        th = threading.Thread(target=reader, name="read_vis-reader", daemon=True)
        th.start()

        def upload(i):
            s = i % depth
            host_full[s].acquire()               # block i has been decoded into pinned set s
            if failure:
                raise failure[0]
            n = blocks[i].stop - blocks[i].start
            with torch.cuda.stream(copy_stream):
                if consumed[s] is not None:      # the kernels that read device set s (block i - depth) come first
                    copy_stream.wait_event(consumed[s])
                for k in names:
                    on_dev[s][k][:n].copy_(pinned[s][k][:n], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            copied[s] = ev

        try:
            upload(0)
            for i, sl in enumerate(blocks):
                s = i % depth
                n = sl.stop - sl.start
                torch.cuda.current_stream(dev).wait_event(copied[s])
                yield sl, {k: on_dev[s][k][:n] for k in names}
                ev = torch.cuda.Event()          # the caller has queued its kernels on block i
                ev.record(torch.cuda.current_stream(dev))
                consumed[s] = ev
                copied[s].synchronize()          # pinned set s is free for the reader once its copy has landed
                host_free[s].release()
                if i + 1 < len(blocks):
                    upload(i + 1)                # waits for the reader; block i's kernels run meanwhile
        finally:
            stop.set()
            for sem in host_free:                # unblock the reader if the caller stopped early
                sem.release()
            th.join(timeout=60)
            torch.cuda.current_stream(dev).wait_stream(copy_stream)


class Mxds:
    """Dataset of datasets: `.attrs[name]` is a VisDataset (vis_xds_packager, cngi/_utils/_io.py:37-40)."""

    def __init__(self, parts):
        self.attrs = dict(parts)

    def copy(self):
        return Mxds(self.attrs)

    def __getattr__(self, name):
        try:
            return self.__dict__["attrs"][name]
        except KeyError:
            raise AttributeError(name)


def read_vis(infile, partition=None, chunks=None, consolidated=True, overwrite_encoded_chunks=True, **kwargs):
    """cngi/dio/read_vis.py:21.  `chunks` ({'time': n, ...}) overrides the on-disk chunking for the chunk walk;
    `consolidated` / `overwrite_encoded_chunks` are accepted for signature parity (metadata is read per array)."""
    if str(infile).lower().startswith("s3"):
        raise NotImplementedError("the S3 branch of read_vis (read_vis.py:63-180) is a network path and is not built")
    infile = os.path.expanduser(infile)
    if partition is None:
        partition = sorted(os.listdir(infile))
    partition = [str(p) for p in np.atleast_1d(partition)]
    if "global" in partition and os.path.isdir(os.path.join(infile, "global")):
        partition += sorted("global/" + t for t in os.listdir(os.path.join(infile, "global")))
    parts = []
    for part in partition:
        if part == "global" or part.startswith("."):
            continue
        path = os.path.join(infile, part)
        if os.path.isdir(path):
            try:
                parts.append((part.replace("global/", ""), VisDataset(path, chunks=chunks)))
            except Exception:
                print("Can not open ", part)            # read_vis.py:196
    return Mxds(parts)


def write_vis(outfile, xds, partition="xds0", chunks=None, compressor=DEFAULT_COMPRESSOR, attrs=None):
    """Writes {name: array} as `<outfile>/<partition>/<name>/` zarr v2 arrays with xarray's `_ARRAY_DIMENSIONS`.

    4-d variables get dims (time, baseline, chan, pol), UVW (time, baseline, uvw_index), FIELD_ID-like 2-d variables
    (time, baseline), `chan` / `time` / ... 1-d coordinates their own name.  chunks = {'time': n, 'baseline': n, ...}.
    """
    chunks = dict(chunks or {})
    root = os.path.join(outfile, partition)
    os.makedirs(root, exist_ok=True)
    for d in (outfile, root):
        with open(os.path.join(d, ".zgroup"), "w") as f:
            json.dump({"zarr_format": 2}, f)
    with open(os.path.join(root, ".zattrs"), "w") as f:
        json.dump(dict(attrs or {}), f)
    for name, arr in xds.items():
        arr = np.asarray(arr)
        if arr.ndim == 4:
            dims = SAMPLE_DIMS
        elif arr.ndim == 3:
            dims = ("time", "baseline", "uvw_index")
        elif arr.ndim == 2:
            dims = ("time", "baseline")
        elif arr.ndim == 1:
            dims = (name,)
        else:
            dims = ()
        ch = tuple(int(chunks.get(d, n)) or n for d, n in zip(dims, arr.shape))
        zs.write_array(os.path.join(root, name), arr, chunks=ch, compressor=compressor, dims=dims)
    return outfile
