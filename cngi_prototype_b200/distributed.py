"""Multi-GPU decomposition of the gridding path: one process per GPU, torch.distributed (NCCL over NVLink).

Replaces the reference's dask decomposition (one delayed task per (time, baseline, chan) chunk, each returning a
full grid, summed by a pairwise `da.add` tree -- /root/reference/ngcasa/imaging/_imaging_utils/_standard_grid.py:58-98,
_tree_sum_list :109-120) with device-resident partial grids and one collective per grid:

  continuum (all channels -> one image plane): samples are sharded along TIME; every rank grids its shard into a
      full partial grid; partial grids + sum_weight are summed with reduce (to the rank that runs the FFT) or
      all-reduce (imaging-weight density: every rank needs all of it for its own degrid).
  cube (channel c -> image plane c): samples are sharded along CHANNEL; each rank owns its image planes end to end
      (grid -> FFT -> normalise), exactly like synthesis_imaging_cube.py:105-124; no exchange is needed.

The collectives are size-independent of the sample count (one uv-grid), so they are issued once per step, after all
local samples are gridded.  The compute operators are injected (`ops`), so the same control flow runs on GPUs with
the CUDA operators (bench.py, NCCL) and on CPUs with the oracle (tests, gloo).
"""
from types import SimpleNamespace

import torch
import torch.distributed as dist


def init_nccl(device):
    """torch.distributed over NCCL with the collectives on a HIGH-PRIORITY stream.

    The collectives of a step are issued while the gridder's ~10 000 blocks are being dispatched.  On a default-priority
    stream the NCCL kernel's blocks queue behind the gridder's pending blocks and the "overlapped" reduce in fact runs
    after it; on a high-priority stream they take the next free SM slots.  Measured on 8 B200s (tools/probe_collectives.py,
    weak scaling, C2): 2.96 -> 2.79 ms per step (standalone: all-reduce of the 134 MB density 0.40 ms, reduce of the
    268 MB grid 0.45 ms; a single GPU takes 2.50 ms)."""
    if dist.is_initialized():
        return
    opts = None
    try:
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
    except Exception:   # older / different builds: fall back to the default stream priority
        opts = None
    if opts is not None:
        dist.init_process_group("nccl", device_id=device, pg_options=opts)
    else:
        dist.init_process_group("nccl", device_id=device)


class _EventWork:
    """The `.wait()` of an async collective for work that was queued on a side stream: the current stream waits for it."""

    def __init__(self, event):
        self.event = event

    def wait(self):
        torch.cuda.current_stream().wait_event(self.event)


class SymmetricCollectives:
    """The two sums of the continuum step done INSIDE THE NVSWITCH (NVLS) by this library's own kernels instead of NCCL.

    The accumulators are allocated in symmetric memory (`torch.distributed._symmetric_memory`: the same buffer on every
    rank, mapped through one multicast address).  `reduce_grid` then costs ONE kernel on the root --
    cngi_b200_multimem_reduce_f32: `multimem.ld_reduce` streams the sum over all ranks out of the switch into the root's
    buffer -- and nothing at all on the other ranks (their HBM is read over NVLink; NCCL runs 24-32 blocks of 640 threads on
    every rank).  `allreduce_density` reduces each rank's 1/N slice of pol plane 0 with `multimem.ld_reduce` and broadcasts
    it with `multimem.st` (cngi_b200_multimem_allreduce_f64).  Ordering across ranks comes from the symmetric-memory barrier
    (device side, on the collective stream): one before (every rank's partial result is complete) and one after (nobody
    overwrites a buffer that is still being read).  Everything runs on one high-priority side stream in program order, so
    the sums overlap the next kernels of the compute stream exactly like the NCCL calls they replace.

    Measured (tools/probe_collectives.py; results identical to NCCL's to 1e-7 / bit for bit): on 8 B200s the all-reduce of
    the 134 MB density takes 0.31 ms (NCCL 0.40 ms), the reduce of the 268 MB grid 0.72 ms (NCCL 0.45 ms: the root has to pull
    the whole grid through its own NVLink port, where NCCL's ring spreads the summation over every rank's SMs and links), and
    the pipelined step 2.85 ms against 2.76 ms with NCCL on a high-priority stream -- so NCCL stays the default and this is
    opt-in (ContinuumPipeline(symmetric=...)); on 2 GPUs there is no switch-side advantage at all (2.78 vs 2.68 ms)."""

    def __init__(self, device, group=None, n_blocks=0):
        import torch.distributed._symmetric_memory as symm_mem
        self.symm_mem = symm_mem
        self.device = device
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.stream = torch.cuda.Stream(device=device, priority=-1)
        self.n_blocks = int(n_blocks)
        self.handles = {}

    @staticmethod
    def supported(device):
        try:
            import torch.distributed._symmetric_memory as symm_mem
            from torch._C._autograd import DeviceType
            return bool(symm_mem._SymmetricMemory.has_multicast_support(DeviceType.CUDA, torch.device(device).index or 0))
        except Exception:
            return False

    def empty(self, shape, dtype):
        """A symmetric tensor (collective: every rank calls it with the same shape, in the same order)."""
        t = self.symm_mem.empty(tuple(int(x) for x in shape), dtype=dtype, device=self.device)
        h = self.symm_mem.rendezvous(t, self.group)
        assert h.multicast_ptr != 0, "no multicast mapping (NVLS) for this buffer"
        self.handles[t.data_ptr()] = h
        return t

    def _bracket(self, t, body):
        """barrier -> body -> barrier on the collective stream, ordered behind the work queued so far on the current
        stream; returns the work handle whose .wait() orders the current stream behind it."""
        from . import _lib
        h = self.handles[t.data_ptr()]
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            h.barrier(channel=0)
            body(h, _lib)
            h.barrier(channel=1)
            done = torch.cuda.Event()
            done.record(self.stream)
        t.record_stream(self.stream)
        return _EventWork(done)

    def reduce_grid(self, grid, root):
        """grid (symmetric, complex64 or float32, contiguous): the sum over ranks lands in the root's buffer."""
        n_floats = grid.numel() * (2 if grid.is_complex() else 1)
        assert grid.is_contiguous() and grid.element_size() * grid.numel() == 4 * n_floats

        def body(h, _lib):
            if self.rank == root:
                _lib.check(_lib.lib().cngi_b200_multimem_reduce_f32(h.multicast_ptr, grid.data_ptr(), n_floats, self.n_blocks,
                                                                    self.stream.cuda_stream), "cngi_b200_multimem_reduce_f32")
        return self._bracket(grid, body)

    def allreduce_density(self, density, n_doubles):
        """the first n_doubles of `density` (symmetric, float64: pol plane 0 of a continuum density) summed on every rank"""
        assert density.dtype == torch.float64 and density.is_contiguous()

        def body(h, _lib):
            _lib.check(_lib.lib().cngi_b200_multimem_allreduce_f64(h.multicast_ptr, int(n_doubles), self.rank, self.world,
                                                                   self.n_blocks, self.stream.cuda_stream),
                       "cngi_b200_multimem_allreduce_f64")
        return self._bracket(density, body)


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank, world_size):
    """Contiguous, balanced [lo, hi) block of n items for `rank` (first n % world_size ranks get one extra)."""
    base, extra = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def time_shard(arrays, rank, world_size):
    """Slices the time axis (axis 0) of every sample array in the dict; freq_chan and scalars pass through."""
    n_time = arrays["uvw"].shape[0]
    lo, hi = shard_range(n_time, rank, world_size)
    out = {}
    for k, v in arrays.items():
        out[k] = v[lo:hi] if (hasattr(v, "shape") and len(v.shape) >= 2 and v.shape[0] == n_time) else v
    return out


def channel_shard(arrays, rank, world_size):
    """Slices the channel axis: axis 2 of the 4-D sample arrays and axis 0 of freq_chan; uvw is replicated."""
    n_chan = arrays["freq_chan"].shape[0]
    lo, hi = shard_range(n_chan, rank, world_size)
    out = {}
    for k, v in arrays.items():
        if k == "freq_chan":
            out[k] = v[lo:hi]
        elif hasattr(v, "shape") and len(v.shape) == 4:
            out[k] = v[:, :, lo:hi]
        else:
            out[k] = v
    return out


def _as_real(t):
    return torch.view_as_real(t) if t.is_complex() else t


def allreduce_sum(*tensors):
    """In-place sum over ranks (complex tensors are reduced as interleaved reals)."""
    if world()[1] > 1:
        for t in tensors:
            dist.all_reduce(_as_real(t))


def reduce_sum(dst, *tensors):
    """In-place sum onto rank `dst` (other ranks' buffers are left unspecified, as with ncclReduce)."""
    if world()[1] > 1:
        for t in tensors:
            dist.reduce(_as_real(t), dst)


def continuum_imaging_step(ops, d, gp, gp_iw, iw_parms, cgk, bufs, grid_hook=None):
    """One time-sharded continuum step on this rank's shard `d` (dict of vis, uvw, weight, freq_chan):

        density grid -> ALL-REDUCE(density, sum_weight) -> Briggs factors -> weight degrid
        -> standard gridding of vis * imaging weight -> REDUCE(grid, sum_weight) to rank 0.

    `ops` provides imaging_weight_grid(uvw, w, freq, gp_iw, grid=, sum_weight=), briggs(density, sw, parms),
    degrid(density, uvw, w, briggs, freq, gp_iw) and standard_grid(vis, uvw, w, freq, cgk, gp, grid=, sum_weight=);
    `bufs` holds the accumulators density, dsw, grid, gsw (zeroed here).  Returns the imaging weights.
    """
    for t in (bufs.density, bufs.dsw, bufs.grid, bufs.gsw):
        t.zero_()
    ops.imaging_weight_grid(d["uvw"], d["weight"], d["freq_chan"], gp_iw, grid=bufs.density, sum_weight=bufs.dsw)
    allreduce_sum(bufs.density, bufs.dsw)          # every rank needs the full density for its own degrid
    bf = ops.briggs(bufs.density, bufs.dsw, iw_parms)
    iw = ops.degrid(bufs.density, d["uvw"], d["weight"], bf, d["freq_chan"], gp_iw)
    if grid_hook is not None:
        grid_hook("begin")
    ops.standard_grid(d["vis"], d["uvw"], iw, d["freq_chan"], cgk, gp, grid=bufs.grid, sum_weight=bufs.gsw)
    if grid_hook is not None:
        grid_hook("end")
    reduce_sum(0, bufs.grid, bufs.gsw)             # partial uv-grids -> the rank that runs the FFT
    return iw


class ContinuumPipeline:
    """Software-pipelined sequence of time-sharded continuum steps (one per chunk / dataset).

    A step is  A: density grid -> all-reduce(density)      B: Briggs -> weight degrid -> gridding -> reduce(grid).
    Both collectives have a consumer right behind them, so inside ONE step they cannot be hidden; across steps they
    can: step k+1's stage A is issued before step k's stage B, so all-reduce(k+1) runs under gridding(k), and
    reduce(k) runs under A(k+2)/B(k+1).  Accumulators are double-buffered; a buffer is reused only after the
    collective that reads it has finished.  With world_size 1 the same kernels run in the same order, minus the
    collectives.  The density all-reduce moves pol plane 0 only (all pol planes are identical when n_pol >= 2).

        pipe = ContinuumPipeline(ops, gp, gp_iw, iw_parms, cgk, make_bufs)
        for d in chunks: pipe.step(d)
        pipe.flush()            # results of the last step: pipe.last (grid, gsw valid on rank 0)
    """

    def __init__(self, ops, gp, gp_iw, iw_parms, cgk, make_bufs, side_stream=None, fuse_weights=False, grid_reduce="root0",
                 symmetric=None):
        """side_stream (a high-priority CUDA stream, device tensors only): the whole imaging-weight chain of step k+1
        (density grid, all-reduce, Briggs factors, weight degrid -- memory-latency bound kernels) is issued there and
        runs CONCURRENTLY with the gridding kernel of step k on the current stream (issue bound): whenever a gridder
        block retires, a pending block of the high-priority stream takes its slot, and the two kinds of warps fill each
        other's stalls."""
        self.ops, self.gp, self.gp_iw, self.iw_parms, self.cgk = ops, gp, gp_iw, iw_parms, cgk
        self.side = side_stream
        # grid_reduce: where the summed uv-grid of a step ends up.  "root0": rank 0 (reduce); "rotate": rank k mod N for
        # step k (successive datasets are transformed by successive ranks: the FFTs and the root's share of the reduce are
        # spread over the GPUs); "allreduce": everywhere; "none": nowhere (measurement only)
        assert grid_reduce in ("root0", "rotate", "allreduce", "none"), grid_reduce
        self.grid_reduce = grid_reduce
        self.last_root = 0
        # symmetric: a SymmetricCollectives whose .empty() allocated the density / grid accumulators of make_bufs: the two
        # big sums then go through the NVSwitch (multimem kernels of this library) instead of NCCL
        self.symmetric = symmetric
        # fuse_weights: the weight degrid (A4) runs inside the gridder (ops.standard_grid_weighted) when the ops have it and
        # the support is 7 -- the imaging weights are then never written or re-read.  Measured on C2 (B200, fp32): the
        # gridder is issue bound, so the folded-in work costs what the separate pass costs (2.53 vs 2.50 ms per step): off
        # by default here, used by the host-array API where it saves a 0.46 GB intermediate (imaging._grid)
        self.fuse_weights = bool(fuse_weights) and int(gp.get("support", 7)) == 7 and hasattr(ops, "standard_grid_weighted")
        self.bufs = [make_bufs(), make_bufs()]
        self.pend_density = [[], []]
        self.pend_grid = [[], []]
        self.k = 0
        self.n_grids = 0
        self.prev = None
        self.last = None

    @staticmethod
    def _wait(works):
        for w in works:
            w.wait()
        del works[:]

    def _stage_a(self, d, slot):
        b = self.bufs[slot]
        n_pol = b.density.shape[1]
        (b.density[:, :1] if n_pol >= 2 else b.density).zero_()   # only plane 0 is gridded when n_pol >= 2
        b.dsw.zero_()
        self.ops.imaging_weight_grid(d["uvw"], d["weight"], d["freq_chan"], self.gp_iw, grid=b.density, sum_weight=b.dsw,
                                     first_pol_only=n_pol >= 2)
        if world()[1] > 1:   # every rank needs the full density for its own degrid; plane 0 carries all the information
            first = b.density[:, :1] if n_pol >= 2 else b.density
            if self.symmetric is not None and b.density.shape[0] == 1:   # continuum: plane 0 is the head of the buffer
                self.pend_density[slot] = [self.symmetric.allreduce_density(b.density, first.numel()),
                                           dist.all_reduce(b.dsw, async_op=True)]
            else:
                self.pend_density[slot] = [dist.all_reduce(first, async_op=True), dist.all_reduce(b.dsw, async_op=True)]

    def _weights(self, d, b):
        """Briggs factors + weight degrid from the (all-reduced) density.  With n_pol >= 2 only pol plane 0 was gridded
        (all planes are identical by construction): the other planes are stride-0 views of it -- no replication pass,
        and the sum of squares reads one plane."""
        n_pol = b.density.shape[1]
        if n_pol >= 2:
            rho0, sw0 = b.density[:, :1], b.dsw[:, :1]
            bf = self.ops.briggs(rho0, sw0, self.iw_parms).expand(-1, -1, n_pol)
            rho = rho0.expand(-1, n_pol, -1, -1)
        else:
            bf, rho = self.ops.briggs(b.density, b.dsw, self.iw_parms), b.density
        return self.ops.degrid(rho, d["uvw"], d["weight"], bf, d["freq_chan"], self.gp_iw)

    def _weight_source(self, b):
        """The fused form of _weights: Briggs factors + what the gridder needs to form the imaging weights itself."""
        n_pol = b.density.shape[1]
        if n_pol >= 2:
            rho0, sw0 = b.density[:, :1], b.dsw[:, :1]
            bf = self.ops.briggs(rho0, sw0, self.iw_parms).expand(-1, -1, n_pol)
            rho = rho0.expand(-1, n_pol, -1, -1)
        else:
            bf, rho = self.ops.briggs(b.density, b.dsw, self.iw_parms), b.density
        return dict(density=rho, briggs_factors=bf, grid_parms=self.gp_iw)

    def _stage_b(self, d, slot, grid_hook):
        b = self.bufs[slot]
        self._wait(self.pend_density[slot])
        fused = self.fuse_weights
        iw = self._weight_source(b) if fused else self._weights(d, b)
        self._wait(self.pend_grid[slot])   # the reduce that last read this grid buffer
        b.grid.zero_()
        b.gsw.zero_()
        if grid_hook is not None:
            grid_hook("begin")
        if fused:   # imaging weights are formed inside the gridder from the natural weights + density (never materialised)
            self.ops.standard_grid_weighted(d["vis"], d["uvw"], d["weight"], d["freq_chan"], self.cgk, self.gp, iw,
                                            grid=b.grid, sum_weight=b.gsw)
        else:
            self.ops.standard_grid(d["vis"], d["uvw"], iw, d["freq_chan"], self.cgk, self.gp, grid=b.grid, sum_weight=b.gsw)
        if grid_hook is not None:
            grid_hook("end")
        self._reduce_grid(b, slot)
        self.last = b
        return None if fused else iw

    def _reduce_grid(self, b, slot):
        """partial uv-grids -> the rank that runs the FFT of this step (self.last_root)"""
        ws = world()[1]
        if ws <= 1 or self.grid_reduce == "none":
            return
        if self.grid_reduce == "allreduce":
            self.pend_grid[slot] = [dist.all_reduce(_as_real(b.grid), async_op=True), dist.all_reduce(b.gsw, async_op=True)]
            return
        root = (self.n_grids % ws) if self.grid_reduce == "rotate" else 0
        self.n_grids += 1
        self.last_root = root
        if self.symmetric is not None and b.grid.dtype in (torch.complex64, torch.float32):
            self.pend_grid[slot] = [self.symmetric.reduce_grid(b.grid, root), dist.reduce(b.gsw, root, async_op=True)]
        else:
            self.pend_grid[slot] = [dist.reduce(_as_real(b.grid), root, async_op=True), dist.reduce(b.gsw, root, async_op=True)]

    def _weights_on_side(self, d, slot):
        """The whole weight chain of a step on the side stream; returns (imaging weights, event that marks them ready)."""
        main = torch.cuda.current_stream()
        self.side.wait_stream(main)   # inputs may have been produced on the current stream
        with torch.cuda.stream(self.side):
            self._stage_a(d, slot)
            b = self.bufs[slot]
            self._wait(self.pend_density[slot])
            iw = self._weight_source(b) if self.fuse_weights else self._weights(d, b)
            ev = torch.cuda.Event()
            ev.record(self.side)
        for t in (iw.values() if isinstance(iw, dict) else (iw,)):
            if torch.is_tensor(t):
                t.record_stream(main)
        return iw, ev

    def _grid_on_main(self, d, slot, grid_hook, iw, ev):
        b = self.bufs[slot]
        torch.cuda.current_stream().wait_event(ev)
        self._wait(self.pend_grid[slot])   # the reduce that last read this grid buffer
        b.grid.zero_()
        b.gsw.zero_()
        if grid_hook is not None:
            grid_hook("begin")
        if isinstance(iw, dict):
            self.ops.standard_grid_weighted(d["vis"], d["uvw"], d["weight"], d["freq_chan"], self.cgk, self.gp, iw,
                                            grid=b.grid, sum_weight=b.gsw)
        else:
            self.ops.standard_grid(d["vis"], d["uvw"], iw, d["freq_chan"], self.cgk, self.gp, grid=b.grid, sum_weight=b.gsw)
        if grid_hook is not None:
            grid_hook("end")
        self._reduce_grid(b, slot)
        self.last = b
        return None if isinstance(iw, dict) else iw

    def step(self, d, grid_hook=None):
        slot = self.k & 1
        if self.side is not None:
            iw, ev = self._weights_on_side(d, slot)
            if self.prev is not None:
                self._grid_on_main(*self.prev)
            self.prev = (d, slot, grid_hook, iw, ev)
        else:
            self._stage_a(d, slot)
            if self.prev is not None:
                self._stage_b(*self.prev)
            self.prev = (d, slot, grid_hook)
        self.k += 1

    def flush(self):
        iw = None
        if self.prev is not None:
            iw = self._grid_on_main(*self.prev) if self.side is not None else self._stage_b(*self.prev)
            self.prev = None
        for slot in (0, 1):
            self._wait(self.pend_density[slot])
            self._wait(self.pend_grid[slot])
        return iw


def make_imaging_weight(vis_dataset_shard, imaging_weights_parms, grid_parms, time_chunk=0):
    """imaging.make_imaging_weight on this rank's TIME SHARD of a continuum dataset: the density grid (pol plane 0: all
    planes are identical) and its sum_weight are all-reduced before the Briggs factors are taken, so every rank degrids
    its samples from the density of the whole observation (make_imaging_weight.py:144-247 over all dask chunks)."""
    from . import imaging
    return imaging.make_imaging_weight(vis_dataset_shard, imaging_weights_parms, grid_parms, time_chunk,
                                       _density_hook=lambda rho, sw: allreduce_sum(rho, sw))


def make_grid(vis_dataset_shard, grid_parms, time_chunk=0, weight_key="IMAGING_WEIGHT", apply_flags=False, lazy=False, dst=0):
    """imaging.make_grid on this rank's time shard; the partial uv-grids and sum_weight are summed onto rank `dst`
    (the rank that runs the FFT) -- the role of the reference's tree sum over chunks (_standard_grid.py:109-120).  Only
    `dst`'s result is the grid of the whole observation."""
    from . import imaging
    return imaging.make_grid(vis_dataset_shard, grid_parms, time_chunk, weight_key, apply_flags, lazy,
                             _grid_hook=lambda g, sw: reduce_sum(dst, g, sw))


def cube_layout(rank, world_size, time_split=1):
    """Ranks form a (channel group) x (time part) grid, time part fastest: returns (chan_group, n_chan_groups,
    time_part, root_rank_of_the_group).  time_split = 1 is pure channel sharding (no exchange at all)."""
    assert world_size % time_split == 0, "time_split must divide the world size"
    cg = rank // time_split
    return cg, world_size // time_split, rank % time_split, cg * time_split


def make_time_groups(world_size, time_split):
    """One process group per channel group (the ranks that split its samples along time).  Collective: every rank
    must call it, with the same arguments."""
    if time_split <= 1 or world_size <= 1:
        return None
    groups = [dist.new_group(list(range(g * time_split, (g + 1) * time_split))) for g in range(world_size // time_split)]
    return groups


class _PhaseTimer:
    """CUDA events around the phases of cube_imaging (grid / reduce / image), summed per phase by .ms()."""

    def __init__(self):
        self.marks = []

    def mark(self, name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.marks.append((name, ev))

    def ms(self):
        """{phase: milliseconds}: the time between a mark and the next one is attributed to the LATER mark's name."""
        torch.cuda.synchronize()
        out = {}
        for (_, e0), (name, e1) in zip(self.marks[:-1], self.marks[1:]):
            if name != "begin":
                out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
        return out


def cube_imaging(ops, d, gp, cgk, chan_chunk=0, time_split=1, groups=None, image_out=None, with_psf=False,
                 presharded=False, timer=None, keep_image=True, rotate_roots=False, overlap=False):
    """Channel-sharded cube imaging with bounded memory (BASELINE config 5; synthesis_imaging_cube.py:105-124,171-220).

    Every rank holds (or can slice) the full sample arrays `d` = {vis, uvw, weight, freq_chan[, flag]}.  The image
    channels are split in contiguous blocks over the channel groups (cube_layout); a group's root owns its planes
    end to end: for each chunk of `chan_chunk` channels (0 = the whole block) the group grids the chunk into ONE
    reusable grid buffer, the `time_split` ranks of the group sum their partial grids + sum_weight onto the root
    (the NCCL grid reduce; skipped when time_split == 1: no exchange is needed then), and the root transforms,
    crops and corrects the chunk (ops.to_image) into its slab of the output cube.  Peak memory is one chunk of
    padded grids, not the cube.

    ops.standard_grid(vis, uvw, w, freq, cgk, gp, grid=, sum_weight=[, flag=]) accumulates into the buffers;
    ops.zeros(shape, complex) allocates; ops.to_image(grid, sum_weight, gp) returns (l, m, chan, pol).
    Returns (image (l, m, n_chan_owned, pol) or None on non-root ranks, sum_weight (n_chan_owned, pol) or None,
    (chan_lo, chan_hi)).

    with_psf: also make the PSF cube of the same samples, the way _synthesis_imaging_cube_std_chunk does
    (synthesis_imaging_cube.py:195-211) -- through ops.grid_image_psf (ONE fused pass over the chunk's samples filling
    both grids) when the ops provide it, else through a second ops.standard_grid_psf pass; the psf grid goes through
    the same reduce and transform.  Returns (image, sum_weight, psf, psf_sum_weight, (chan_lo, chan_hi)) then.

    presharded: `d` already holds ONLY this rank's share -- the channels of its channel group and the integrations of
    its time part (how a rank that reads its own zarr chunks, or bench.py, holds the data); nothing is sliced then.
    timer: a _PhaseTimer; CUDA events are recorded after each chunk's gridding ("grid"), reduce ("reduce") and
    transform ("image").  keep_image=False drops each chunk's image after computing it (benchmarks of cubes whose
    output would not fit next to the samples); sum_weight is still returned.
    rotate_roots (time_split > 1): the root of a chunk rotates over the ranks of the group (chunk j -> member j mod
    time_split) and the group walks `time_split` chunks per round through as many grid buffers: after the reduces every
    member transforms ITS chunk, so the FFTs -- 90 % of a config-5 step -- run on all GPUs instead of on one per group.
    Each rank then returns the planes it owns: the last element of the result is the list of their (chan_lo, chan_hi).
    overlap (time_split == 1, CUDA): grid chunk j + 1 on the current stream while chunk j is transformed on a side stream
    (two grid buffers); the per-phase times of `timer` then overlap and only their sum against the step time is meaningful.
    """
    assert keep_image or not with_psf, "keep_image=False is a benchmark mode of the image-only path"
    rank, ws = world()
    cg, n_cg, tp, root = cube_layout(rank, ws, time_split)
    n_time, n_chan = d["uvw"].shape[0], d["freq_chan"].shape[0]
    if presharded:
        (clo, chi), (tlo, thi) = (0, n_chan), (0, n_time)
    else:
        clo, chi = shard_range(n_chan, cg, n_cg)
        tlo, thi = shard_range(n_time, tp, time_split)
    n_pol = d["weight"].shape[3]
    n_u, n_v = (int(x) for x in gp["image_size_padded"])
    step = int(chan_chunk) if chan_chunk else max(chi - clo, 1)
    gpc = dict(gp, chan_mode="cube")
    group = groups[cg] if groups else None
    rotate = bool(rotate_roots) and group is not None and time_split > 1
    n_buf = time_split if rotate else 1
    n_planes = min(step, max(chi - clo, 1))

    def make_buf():
        b = SimpleNamespace(grid=ops.zeros((n_planes, n_pol, n_u, n_v), True), gsw=ops.zeros((n_planes, n_pol), False),
                            pgrid=None, pgsw=None)
        if with_psf:
            b.pgrid = torch.zeros(tuple(b.grid.shape), dtype=b.grid.real.dtype, device=b.grid.device)
            b.pgsw = ops.zeros((n_planes, n_pol), False)
        return b

    bufs = [make_buf() for _ in range(n_buf)]
    out = SimpleNamespace(image=image_out, sum_weight=None, psf=None, psf_sum_weight=None, owned=[])

    def mark(name):
        if timer is not None:
            timer.mark(name)

    def grid_chunk(c0, c1, b):
        g, s = b.grid[:c1 - c0], b.gsw[:c1 - c0]
        g.zero_()
        s.zero_()
        kw = {}
        if d.get("flag") is not None:
            kw["flag"] = d["flag"][tlo:thi, :, c0:c1]
        if with_psf:
            b.pgrid[:c1 - c0].zero_()
            b.pgsw[:c1 - c0].zero_()
        if thi > tlo:
            args = (d["uvw"][tlo:thi], d["weight"][tlo:thi, :, c0:c1], d["freq_chan"][c0:c1], cgk, gpc)
            if with_psf and hasattr(ops, "grid_image_psf"):
                ops.grid_image_psf(d["vis"][tlo:thi, :, c0:c1], *args, grid=g, sum_weight=s, psf_grid=b.pgrid[:c1 - c0],
                                   psf_sum_weight=b.pgsw[:c1 - c0], **kw)
            else:
                ops.standard_grid(d["vis"][tlo:thi, :, c0:c1], *args, grid=g, sum_weight=s, **kw)
                if with_psf:
                    ops.standard_grid_psf(*args, grid=b.pgrid[:c1 - c0], sum_weight=b.pgsw[:c1 - c0])

    def reduce_chunk(c0, c1, b, dst):
        if group is None:
            return
        dist.reduce(_as_real(b.grid[:c1 - c0]), dst, group=group)
        dist.reduce(b.gsw[:c1 - c0], dst, group=group)
        if with_psf:
            dist.reduce(b.pgrid[:c1 - c0], dst, group=group)
            dist.reduce(b.pgsw[:c1 - c0], dst, group=group)

    def image_chunk(c0, c1, b):
        g, s = b.grid[:c1 - c0], b.gsw[:c1 - c0]
        img = ops.to_image(g, s, gpc)
        out.owned.append((c0, c1))
        if out.sum_weight is None:
            out.sum_weight = ops.zeros((chi - clo, n_pol), False)
        out.sum_weight[c0 - clo:c1 - clo] = s
        if not keep_image:
            return
        if out.image is None:
            out.image = ops.zeros(tuple(img.shape[:2]) + (chi - clo, n_pol), False).to(img.dtype)
        out.image[:, :, c0 - clo:c1 - clo] = img
        if with_psf:
            pimg = ops.to_image(b.pgrid[:c1 - c0], b.pgsw[:c1 - c0], gpc)
            if out.psf is None:
                out.psf = ops.zeros(tuple(pimg.shape[:2]) + (chi - clo, n_pol), False).to(pimg.dtype)
                out.psf_sum_weight = ops.zeros((chi - clo, n_pol), False)
            out.psf[:, :, c0 - clo:c1 - clo] = pimg
            out.psf_sum_weight[c0 - clo:c1 - clo] = b.pgsw[:c1 - c0]

    chunks = [(c0, min(chi, c0 + step)) for c0 in range(clo, chi, step)]
    if overlap and group is None and bufs[0].grid.is_cuda and len(chunks) > 1:
        # Pure channel sharding on a GPU: the gridding of chunk j + 1 (bound by the reductions into a chunk of planes far
        # larger than L2: ~30 % of the issue slots) runs on the current stream while the FFTs of chunk j (shared-memory bound)
        # run on a side stream -- two grid buffers, each re-zeroed only after its transform has finished.
        dev = bufs[0].grid.device
        side = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        bufs.append(make_buf())
        free = [None, None]
        for j, (c0, c1) in enumerate(chunks):
            b = bufs[j & 1]
            if free[j & 1] is not None:
                main.wait_event(free[j & 1])
            mark("begin")
            grid_chunk(c0, c1, b)
            mark("grid")
            gridded = torch.cuda.Event()
            gridded.record(main)
            with torch.cuda.stream(side):
                side.wait_event(gridded)
                image_chunk(c0, c1, b)
                free[j & 1] = torch.cuda.Event()
                free[j & 1].record(side)
        main.wait_stream(side)
        mark("image")
        for t in (out.image, out.sum_weight, out.psf, out.psf_sum_weight):
            if t is not None:   # allocated under the side stream, handed to the caller's stream
                t.record_stream(main)
    else:
        for j0 in range(0, len(chunks), n_buf):
            todo = chunks[j0:j0 + n_buf]
            mark("begin")
            mine = []
            for i, (c0, c1) in enumerate(todo):          # every rank of the group grids its time part of each chunk ...
                grid_chunk(c0, c1, bufs[i])
                mark("grid")
                dst = root + ((j0 + i) % time_split if rotate else 0)
                reduce_chunk(c0, c1, bufs[i], dst)       # ... and the partial grids are summed onto that chunk's root
                if group is not None:
                    mark("reduce")
                if rank == dst:
                    mine.append((c0, c1, bufs[i]))
            for c0, c1, b in mine:                        # the roots transform their chunks at the same time
                image_chunk(c0, c1, b)
                mark("image")
    owner = bool(out.owned) or rank == root
    rng = out.owned if rotate else (clo, chi)
    if with_psf:
        return (out.image, out.sum_weight, out.psf, out.psf_sum_weight, rng) if owner else (None, None, None, None, rng)
    return (out.image, out.sum_weight, rng) if owner else (None, None, rng)


def cuda_ops():
    """The product operators (libcngi_b200.so through the Python mirror) in the shape continuum_imaging_step expects."""
    from ._standard_grid import standard_grid
    from ._imaging_weight import (imaging_weight_grid, calculate_briggs_parms,
                                  _standard_imaging_weight_degrid_numpy_wrap)

    def degrid(density, uvw, w, bf, freq, gp_iw):
        return _standard_imaging_weight_degrid_numpy_wrap(density, uvw, w, bf, freq, gp_iw, kernel_side_layout=True)

    def grid(vis, uvw, w, freq, cgk, gp, grid=None, sum_weight=None, flag=None):
        return standard_grid(vis, uvw, w, freq, cgk, gp, False, True, grid=grid, sum_weight=sum_weight, flag=flag)

    def grid_weighted(vis, uvw, w_nat, freq, cgk, gp, source, grid=None, sum_weight=None, flag=None):
        return standard_grid(vis, uvw, w_nat, freq, cgk, gp, False, True, grid=grid, sum_weight=sum_weight, flag=flag,
                             imaging_weight_from=source)

    def grid_image_psf(vis, uvw, w, freq, cgk, gp, grid=None, sum_weight=None, psf_grid=None, psf_sum_weight=None, flag=None):
        from ._standard_grid import standard_grid_image_psf
        return standard_grid_image_psf(vis, uvw, w, freq, cgk, gp, flag=flag, grid=grid, sum_weight=sum_weight,
                                       psf_grid=psf_grid, psf_sum_weight=psf_sum_weight)

    def grid_psf(uvw, w, freq, cgk, gp, grid=None, sum_weight=None):
        return standard_grid(None, uvw, w, freq, cgk, gp, True, False, grid=grid, sum_weight=sum_weight)

    def zeros(shape, is_complex, precision="f32"):
        dt = {("f32", True): torch.complex64, ("f32", False): torch.float64,
              ("f64", True): torch.complex128, ("f64", False): torch.float64}[(precision, bool(is_complex))]
        return torch.zeros(shape, dtype=dt, device="cuda")

    def to_image(g, s, gp):
        from ._fft import grid_to_image
        from ._gridding_convolutional_kernels import correcting_function_1D
        cu, cv = correcting_function_1D(gp["image_size_padded"], gp["image_size"])
        return grid_to_image(g, gp["image_size"], sum_weight=s, corr_u=cu, corr_v=cv)

    return SimpleNamespace(imaging_weight_grid=imaging_weight_grid, briggs=calculate_briggs_parms, degrid=degrid,
                           standard_grid=grid, standard_grid_weighted=grid_weighted, standard_grid_psf=grid_psf, grid_image_psf=grid_image_psf, zeros=zeros,
                           to_image=to_image)
