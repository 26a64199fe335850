"""Host-array plumbing of the API layer: per-thread CUDA streams, chunked host->device feeding, lazy device results.

The reference's API functions return LAZY datasets: `make_imaging_weight` adds a dask array to the dataset and
`make_grid` / `make_image` consume it; nothing is computed or moved until `.compute()` / `.values`
(/root/reference/ngcasa/imaging/make_imaging_weight.py:95-104, make_image.py:86-156).  The equivalents here, for
datasets whose variables are HOST (numpy) arrays:

  * `LazyDeviceArray` -- a result that stays on the GPU, ordered behind the kernels that produce it, until somebody asks
    for host values (`np.asarray(x)`, `x[...]`, `x.numpy()`).  `make_imaging_weight` returns IMAGING_WEIGHT as one, so a
    following `make_grid` / `make_image` uses the device copy (and the device copies of UVW / WEIGHT it rode in with)
    and the weights never cross PCIe unless the caller looks at them.
  * `ChunkFeeder` -- walks the time axis of host arrays in chunks: the H2D copy of chunk k+1 (copy stream, straight from
    the caller's memory when it is page-locked) runs while the caller's kernels for chunk k run (compute stream).  Device
    arrays are sliced, not copied.
  * `streams()` -- a (compute, copy, d2h) stream set PER HOST THREAD and device, so that API calls issued from several
    threads at once (how a dask worker runs the `nogil` chunk functions) overlap on the GPU and on both PCIe directions
    instead of serialising on the legacy default stream.

PyTorch is used for device memory, streams and events only.
"""
import threading

import numpy as np

from ._devutil import torch, is_torch

_tls = threading.local()


class _Streams:
    def __init__(self, device):
        self.device = device
        self.compute = torch.cuda.Stream(device=device)
        self.copy = torch.cuda.Stream(device=device)      # host -> device feeding
        self.d2h = torch.cuda.Stream(device=device)       # results -> host (PCIe is full duplex: overlaps the feeding)


def streams(device):
    """The calling thread's stream pair on `device` (created on first use)."""
    table = getattr(_tls, "table", None)
    if table is None:
        table = _tls.table = {}
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    if key not in table:
        table[key] = _Streams(torch.device("cuda", key[1]))
    return table[key]


def is_host(x):
    return not is_torch(x) or not x.is_cuda


def _as_host_tensor(x):
    if is_torch(x):
        return x
    a = np.asarray(x)
    if not a.flags.c_contiguous:
        a = np.ascontiguousarray(a)
    if not a.flags.writeable:      # torch.from_numpy warns on read-only arrays; the tensor is only read
        a = a.view()
        try:
            a.flags.writeable = True
        except ValueError:
            a = a.copy()
    return torch.from_numpy(a)


class LazyDeviceArray:
    """A device-resident result with numpy semantics on demand.

    `base` is the contiguous device tensor, `dims` an optional permutation giving the API-side view (the reference's
    `moveaxis`: a view, not a copy).  `ready` is a CUDA event recorded behind the producing kernels.  The host copy is made
    once, into page-locked memory, in the base layout (one flat DMA), and viewed with the same permutation.
    `sources` carries the device copies of the host inputs the result was computed from -- {name: (host object, device
    tensor)} -- so that the next API call on the same dataset does not upload them again."""

    def __init__(self, base, ready=None, dims=None, sources=None, deferred=None):
        """deferred: the array has not been computed yet -- a dict with `make` (callable returning the contiguous device
        tensor, run on the then-current stream), `shape`, `dtype` (torch), `device`, plus whatever a consumer needs to
        avoid the computation altogether (make_imaging_weight stores the density / Briggs factors / natural weights there,
        and make_grid forms the imaging weights inside the gridder instead of materialising them)."""
        self._base, self.ready, self.dims = base, ready, dims
        self.sources = sources or {}
        self.deferred = deferred
        self._host = None

    @property
    def base(self):
        if self._base is None:
            dev = self.deferred["device"]
            with torch.cuda.device(dev):
                if self.ready is not None:
                    torch.cuda.current_stream(dev).wait_event(self.ready)
                self._base = self.deferred["make"]()
                self.ready = torch.cuda.Event()
                self.ready.record(torch.cuda.current_stream(dev))
        return self._base

    @property
    def computed(self):
        return self._base is not None

    @property
    def device(self):
        return self._base.device if self._base is not None else self.deferred["device"]

    # ---- device side ------------------------------------------------------------------------------------------
    def device_tensor(self):
        """The device tensor (API-side view), safe to use on the CURRENT stream (computed now if it was deferred)."""
        base = self.base
        if self.ready is not None:
            torch.cuda.current_stream(base.device).wait_event(self.ready)
        return base if self.dims is None else base.permute(*self.dims)

    def source(self, name, host_obj):
        """Device copy of `host_obj` if this result was computed from that very object, else None."""
        hit = self.sources.get(name)
        return hit[1] if hit is not None and hit[0] is host_obj else None

    # ---- host side --------------------------------------------------------------------------------------------
    def numpy(self):
        if self._host is None:
            dev = self.device
            s = streams(dev)
            if self._base is None:   # deferred: compute on the thread's compute stream first
                with torch.cuda.device(dev), torch.cuda.stream(s.compute):
                    self.base
            with torch.cuda.device(dev), torch.cuda.stream(s.d2h):
                # its own stream: only `ready` is waited for, so the copy overlaps whatever the caller queued since
                # (e.g. the next dataset's feeding and kernels)
                if self.ready is not None:
                    s.d2h.wait_event(self.ready)
                self.base.record_stream(s.d2h)
                pinned = torch.empty(self.base.shape, dtype=self.base.dtype, pin_memory=True)
                pinned.copy_(self.base, non_blocking=True)
                s.d2h.synchronize()
            host = pinned if self.dims is None else pinned.permute(*self.dims)
            self._host = host.numpy()
        return self._host

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype, copy=False)

    def __getitem__(self, idx):
        return self.numpy()[idx]

    def __len__(self):
        return self.shape[0]

    @property
    def shape(self):
        s = tuple(self._base.shape) if self._base is not None else tuple(self.deferred["shape"])
        return s if self.dims is None else tuple(s[d] for d in self.dims)

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def dtype(self):
        tdt = self._base.dtype if self._base is not None else self.deferred["dtype"]
        return np.dtype(str(tdt).replace("torch.", ""))

    def __repr__(self):
        return "LazyDeviceArray(shape=%s, dtype=%s, device=%s, computed=%s, on_host=%s)" % (
            self.shape, self.dtype, self.device, self.computed, self._host is not None)


def materialise(x):
    """numpy view of a LazyDeviceArray; anything else passes through."""
    return x.numpy() if isinstance(x, LazyDeviceArray) else x


class ChunkFeeder:
    """Time chunks of a set of sample arrays on the device.

        feeder = ChunkFeeder({"UVW": uvw, "WEIGHT": w}, {"UVW": torch.float64, "WEIGHT": None}, device, time_chunk)
        for sl, blk in feeder:            # blk[name]: device tensor of the chunk, ready on the current stream
            kernel(blk["UVW"], blk["WEIGHT"], ...)
        feeder.full("UVW")                # the whole array on the device (all chunks have been fed)

    Host arrays get ONE device buffer; chunk k+1 is copied on the thread's copy stream while the kernels the caller
    queued for chunk k run on the current (compute) stream.  The copy is a direct DMA when the host memory is
    page-locked (torch pinned tensors, or numpy views of them); otherwise the driver stages it.  Device arrays are
    sliced.  `known` maps names to device tensors that are already resident (from a LazyDeviceArray's sources)."""

    def __init__(self, arrays, dtypes, device, time_chunk=0, known=None, min_chunks=8):
        self.device = device
        self.names = list(arrays)
        self.n_time = int(arrays[self.names[0]].shape[0])
        self.dev, self.host = {}, {}
        known = known or {}
        for name, x in arrays.items():
            dt = dtypes.get(name)
            if known.get(name) is not None:
                self.dev[name] = known[name]
            elif is_torch(x) and x.is_cuda:
                self.dev[name] = x if (dt is None or x.dtype == dt) else x.to(dt)
            else:
                h = _as_host_tensor(x)
                if dt is not None and h.dtype != dt:
                    h = h.to(dt)                     # dtype conversion on the host (rare: e.g. bool FLAG -> uint8)
                self.host[name] = h
        step = int(time_chunk) if time_chunk else 0
        if step <= 0:
            step = self.n_time if not self.host else max(1, -(-self.n_time // min_chunks))
        self.slices = [slice(t, min(self.n_time, t + step)) for t in range(0, self.n_time, max(step, 1))]
        self.streams = streams(device) if self.host else None
        if self.host:
            # The device buffers belong to the COPY stream (allocated under it, so the caching allocator orders their reuse
            # against that stream): the feeding of this call can then start while the kernels of the previous call are
            # still running -- a copy.wait_stream(compute) here would insert a bubble of one kernel tail per call.  They
            # are used by the kernels of the compute stream, hence record_stream.
            main = torch.cuda.current_stream(device)
            with torch.cuda.stream(self.streams.copy):
                for name, h in self.host.items():
                    t = torch.empty(h.shape, dtype=h.dtype, device=device)
                    t.record_stream(main)
                    self.dev[name] = t

    def full(self, name):
        return self.dev[name]

    def sources(self, originals):
        """{name: (original host object, device tensor)} for the host arrays that were uploaded."""
        return {name: (originals[name], self.dev[name]) for name in self.host}

    def _copy(self, sl):
        s = self.streams
        with torch.cuda.stream(s.copy):
            for name, h in self.host.items():
                self.dev[name][sl].copy_(h[sl], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s.copy)
        return ev

    def __iter__(self):
        if not self.host:
            for sl in self.slices:
                yield sl, {n: self.dev[n][sl] for n in self.names}
            return
        main = torch.cuda.current_stream(self.device)
        evs = [self._copy(sl) for sl in self.slices]  # the copy queue runs ahead; each chunk's kernels wait for its event
        for sl, ev in zip(self.slices, evs):
            main.wait_event(ev)
            yield sl, {n: self.dev[n][sl] for n in self.names}
