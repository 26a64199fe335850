"""N4 host side: zarr v2 directory store without the zarr package, blosc frames, read_vis / write_vis, chunk walk.

No GPU needed.  The blosc decoder is checked against hand-assembled c-blosc 1.x frames (header + block offsets + split
payloads, the layout documented in cngi_prototype_b200/_zarr_store.py) as well as the module's own writer.
"""
import json
import os
import struct
import zlib

import numpy as np
import pytest

from cngi_prototype_b200 import _zarr_store as zs
from cngi_prototype_b200 import read_vis as rv

COMPRESSORS = [None, {"id": "zlib", "level": 1}, rv.DEFAULT_COMPRESSOR,
               {"id": "blosc", "cname": "zstd", "clevel": 0, "shuffle": 0, "blocksize": 0},      # memcpy frames
               {"id": "blosc", "cname": "lz4", "clevel": 5, "shuffle": 1, "blocksize": 0},
               {"id": "blosc", "cname": "zlib", "clevel": 5, "shuffle": 1, "blocksize": 256}]


def _vis(seed=3, shape=(13, 7, 5, 2)):
    rng = np.random.default_rng(seed)
    vis = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    return {"DATA": vis, "UVW": rng.standard_normal(shape[:2] + (3,)), "WEIGHT": rng.uniform(0.5, 1.5, shape),
            "FLAG": rng.random(shape) < 0.1, "FIELD_ID": rng.integers(0, 3, shape[:2]),
            "chan": np.linspace(1.0e9, 1.1e9, shape[2])}


@pytest.mark.parametrize("comp", COMPRESSORS, ids=lambda c: "raw" if c is None else c["id"] + c.get("cname", "") + str(c.get("clevel", "")))
def test_array_round_trip_ragged_chunks_and_regions(tmp_path, comp):
    a = _vis()["DATA"]
    zs.write_array(str(tmp_path / "DATA"), a, chunks=(4, 7, 2, 2), compressor=comp, dims=rv.SAMPLE_DIMS)
    za = zs.ZarrArray(str(tmp_path / "DATA"))
    assert za.shape == a.shape and za.chunks == (4, 7, 2, 2) and za.dims == rv.SAMPLE_DIMS and za.dtype == a.dtype
    # edge chunks are stored at full chunk shape (zarr v2): 4 time chunks x 3 chan chunks
    assert sorted(f for f in os.listdir(tmp_path / "DATA") if not f.startswith(".")) == \
        sorted("%d.0.%d.0" % (i, j) for i in range(4) for j in range(3))
    assert np.array_equal(np.asarray(za), a)
    assert np.array_equal(za[3:11, :, 1:4], a[3:11, :, 1:4])
    assert np.array_equal(za[-1, 2], a[-1, 2])
    assert np.array_equal(za[..., 1], a[..., 1])
    assert za[5:5].shape == (0, 7, 5, 2)
    with zs.make_pool(4) as pool:
        out = np.empty((9, 7, 5, 2), a.dtype)
        za.read((slice(2, 11),), out=out, pool=pool)
    assert np.array_equal(out, a[2:11])
    with pytest.raises(IndexError):
        za[13]
    with pytest.raises(NotImplementedError):
        za[::2]


def test_missing_chunk_is_fill_value_and_bool_dtype(tmp_path):
    a = np.arange(24, dtype=np.float64).reshape(6, 4)
    zs.write_array(str(tmp_path / "A"), a, chunks=(2, 4), fill_value=float("nan"))
    os.remove(tmp_path / "A" / "1.0")
    got = np.asarray(zs.ZarrArray(str(tmp_path / "A")))
    assert np.array_equal(got[:2], a[:2]) and np.isnan(got[2:4]).all() and np.array_equal(got[4:], a[4:])
    f = np.random.default_rng(0).random((5, 3)) < 0.5
    zs.write_array(str(tmp_path / "F"), f, chunks=(2, 3), compressor=rv.DEFAULT_COMPRESSOR)
    meta = json.load(open(tmp_path / "F" / ".zarray"))
    assert meta["dtype"] == "|b1" and meta["zarr_format"] == 2 and meta["order"] == "C"
    assert np.array_equal(np.asarray(zs.ZarrArray(str(tmp_path / "F"))), f)


def _frame(flags, typesize, raw, blocksize, blocks):
    """c-blosc 1.x frame from already-encoded blocks (each a list of split payloads)."""
    n_blocks = len(blocks)
    pos = 16 + 4 * n_blocks
    bstarts, body = [], b""
    for splits in blocks:
        bstarts.append(pos)
        for p in splits:
            body += struct.pack("<i", len(p)) + p
            pos += 4 + len(p)
    head = struct.pack("<BBBBIII", 2, 1, flags, typesize, len(raw), blocksize, pos)
    return head + struct.pack("<%di" % n_blocks, *bstarts) + body


def test_blosc_hand_assembled_split_blocks_shuffle_and_leftover():
    """typesize 8, blocksize 1024 (= 128 elements, the smallest that splits), byte shuffle, zlib codec (id 3):
    two full blocks of 8 splits each and one unsplit left-over block; one split stored raw (cbytes == split size)."""
    rng = np.random.default_rng(5)
    x = np.cumsum(rng.integers(0, 3, 300)).astype(np.float64)
    raw = x.tobytes()
    blocks = []
    for b in range(3):
        blk = np.frombuffer(raw[b * 1024:(b + 1) * 1024], np.uint8)
        n = len(blk) // 8
        sh = blk.reshape(n, 8).T.reshape(-1).tobytes()               # byte shuffle
        if len(blk) == 1024:
            splits = [sh[k * 128:(k + 1) * 128] for k in range(8)]
            enc = [zlib.compress(s) for s in splits]
            enc[0] = splits[0]                                       # stored: cbytes == 128
            blocks.append(enc)
        else:
            blocks.append([zlib.compress(sh)])                       # the left-over block is never split
    frame = _frame((3 << 5) | 0x1, 8, raw, 1024, blocks)
    assert zs.blosc_decode(frame) == raw
    # do-not-split flag: same blocks, one payload each
    blocks1 = []
    for b in range(3):
        blk = np.frombuffer(raw[b * 1024:(b + 1) * 1024], np.uint8)
        blocks1.append([zlib.compress(blk.reshape(len(blk) // 8, 8).T.reshape(-1).tobytes())])
    assert zs.blosc_decode(_frame((3 << 5) | 0x10 | 0x1, 8, raw, 1024, blocks1)) == raw
    # memcpy frame
    assert zs.blosc_decode(struct.pack("<BBBBIII", 2, 1, 0x2, 8, len(raw), 1024, 16 + len(raw)) + raw) == raw


def test_blosc_rejects_what_it_cannot_read():
    raw = bytes(64)
    good = zs.blosc_encode(raw, 8)
    with pytest.raises(zs.ZarrFormatError):
        zs.blosc_decode(good[:-1])                                   # header cbytes != file size
    with pytest.raises(NotImplementedError):
        zs.blosc_decode(struct.pack("<BBBBIII", 2, 1, 0x4, 8, 64, 64, 16) )       # bit shuffle
    bad = bytearray(good)
    bad[2] = (0 << 5) | 0x10                                         # blosclz
    with pytest.raises(NotImplementedError):
        zs.blosc_decode(bytes(bad))
    with pytest.raises(zs.ZarrFormatError):
        zs.blosc_decode(b"\x03" + good[1:])                          # unknown format version


def test_read_vis_partitions_chunks_and_host_chunk_walk(tmp_path):
    d = _vis()
    store = str(tmp_path / "sim.vis.zarr")
    rv.write_vis(store, d, chunks={"time": 4, "chan": 2})
    rv.write_vis(store, {"NAME_ID": np.arange(3)}, partition="global/FIELD")
    mxds = rv.read_vis(store)
    assert sorted(mxds.attrs) == ["FIELD", "xds0"]                   # default: every partition + global/* (read_vis.py:184-189)
    assert list(rv.read_vis(store, partition="xds0").attrs) == ["xds0"]
    assert sorted(rv.read_vis(store, partition=["xds0", "global"]).attrs) == ["FIELD", "xds0"]
    xds = mxds.xds0
    assert xds.dims == {"time": 13, "baseline": 7, "chan": 5, "pol": 2, "uvw_index": 3}
    assert xds.chunks["time"] == 4 and xds.chunks["chan"] == 2 and xds.chunks["baseline"] == 7
    assert sorted(xds.data_vars) == ["DATA", "FIELD_ID", "FLAG", "UVW", "WEIGHT"]
    for k, v in d.items():
        assert np.array_equal(np.asarray(xds[k]), v), k
    got = list(xds.iter_host_chunks(["DATA", "UVW", "FLAG"], workers=3))
    assert [sl for sl, _ in got] == [slice(0, 4), slice(4, 8), slice(8, 12), slice(12, 13)]
    assert np.array_equal(np.concatenate([b["DATA"] for _, b in got]), d["DATA"])
    assert np.array_equal(np.concatenate([b["FLAG"] for _, b in got]), d["FLAG"])
    assert [sl for sl, _ in xds.iter_host_chunks(["UVW"], time_chunk=6)] == [slice(0, 6), slice(6, 12), slice(12, 13)]
    assert rv.read_vis(store, partition="xds0", chunks={"time": 5}).xds0.chunks["time"] == 5
    with pytest.raises(ValueError):
        list(xds.iter_host_chunks(["chan"]))                         # no leading time axis
    with pytest.raises(NotImplementedError):
        rv.read_vis("s3://bucket/x.vis.zarr")


def test_device_pipeline_and_apply_flags_fail_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cngi_prototype_b200 import _lib, apply_flags as af
    store = rv.write_vis(str(tmp_path / "v.zarr"), _vis())
    xds = rv.read_vis(store).xds0
    with pytest.raises(_lib.CngiError):
        next(iter(xds.iter_device_chunks()))
    with pytest.raises(_lib.CngiError):
        af.apply_flags_chunk(np.zeros(4), np.zeros(4, bool))


def test_oracle_apply_flags_variable_bits(oracle):
    """The restated where().astype(): quiet NaN bit patterns, complex fill NaN + NaN j, unflagged values untouched."""
    rng = np.random.default_rng(1)
    flag = rng.random(64) < 0.3
    for dt, bits, nan in ((np.float32, np.uint32, 0x7fc00000), (np.float64, np.uint64, 0x7ff8000000000000)):
        x = rng.standard_normal(64).astype(dt)
        y = oracle.apply_flags_variable(x, flag)
        assert (y.view(bits)[flag] == nan).all() and np.array_equal(y[~flag], x[~flag]) and y.dtype == dt
        z = (x + 1j * x).astype(np.result_type(dt, np.complex64))
        w = oracle.apply_flags_variable(z, flag)
        assert (w.view(dt).view(bits).reshape(-1, 2)[flag] == nan).all() and np.array_equal(w[~flag], z[~flag])
    with pytest.raises(TypeError):
        oracle.apply_flags_variable(np.arange(4), np.zeros(4, bool))


# ---- native chunk reader (cngi_b200_zarr_read_chunks, host-only code of libcngi_b200.so) ----------------------------
@pytest.mark.parametrize("comp", COMPRESSORS, ids=lambda c: "raw" if c is None else c["id"] + c.get("cname", "") + str(c.get("clevel", "")))
def test_native_reader_equals_python_decoder(tmp_path, comp):
    """Two independent decoders (Python + pyarrow/zlib; C++ + libzstd/libz/own lz4) agree byte for byte on full reads,
    ragged regions, integer indices, several thread counts, every dtype the path stores."""
    d = _vis()
    for name, a in d.items():
        if a.ndim < 2:
            continue
        ch = (4, 7, 2, 2)[:a.ndim] if a.ndim != 3 else (4, 7, 3)
        zs.write_array(str(tmp_path / name), a, chunks=ch, compressor=comp)
        za = zs.ZarrArray(str(tmp_path / name))
        for threads in (1, 3):
            got = za.read(threads=threads)
            assert got.dtype == a.dtype and np.array_equal(got.view(np.uint8), np.ascontiguousarray(a).view(np.uint8)), name
        assert np.array_equal(za.read((slice(3, 11),), threads=2), a[3:11])
        assert np.array_equal(za.read((slice(2, 9), slice(1, 6)), threads=2), a[2:9, 1:6])
        assert np.array_equal(za.read((-1, 2), threads=1), a[-1, 2])
        assert za.read((slice(5, 5),), threads=2).shape == (0,) + a.shape[1:]


def test_native_reader_split_frames_missing_chunks_and_errors(tmp_path):
    from cngi_prototype_b200 import _lib
    # the hand-assembled split / shuffled / stored / left-over frame of the test above, as a chunk file
    rng = np.random.default_rng(5)
    x = np.cumsum(rng.integers(0, 3, 300)).astype(np.float64)
    raw = x.tobytes()
    blocks = []
    for b in range(3):
        blk = np.frombuffer(raw[b * 1024:(b + 1) * 1024], np.uint8)
        sh = blk.reshape(len(blk) // 8, 8).T.reshape(-1).tobytes()
        if len(blk) == 1024:
            enc = [zlib.compress(sh[k * 128:(k + 1) * 128]) for k in range(8)]
            enc[0] = sh[:128]
            blocks.append(enc)
        else:
            blocks.append([zlib.compress(sh)])
    zs.write_array(str(tmp_path / "X"), x, chunks=(300,), compressor={"id": "blosc", "cname": "zlib", "clevel": 5, "shuffle": 1})
    with open(tmp_path / "X" / "0", "wb") as f:
        f.write(_frame((3 << 5) | 0x1, 8, raw, 1024, blocks))
    za = zs.ZarrArray(str(tmp_path / "X"))
    assert np.array_equal(za.read(threads=1), x) and np.array_equal(za.read(), x)
    # lz4 frames with long literal / match runs (length bytes of 255) and overlapping matches
    y = np.zeros(5000, np.int32)
    y[::7] = 3
    y[1000:1300] = np.arange(300)
    zs.write_array(str(tmp_path / "Y"), y, chunks=(5000,), compressor={"id": "blosc", "cname": "lz4", "clevel": 5, "shuffle": 0})
    assert np.array_equal(zs.ZarrArray(str(tmp_path / "Y")).read(threads=1), y)
    # missing chunk -> fill value; NaN fill for floats
    a = np.arange(24, dtype=np.float64).reshape(6, 4)
    zs.write_array(str(tmp_path / "A"), a, chunks=(2, 4), fill_value=float("nan"), compressor=rv.DEFAULT_COMPRESSOR)
    os.remove(tmp_path / "A" / "1.0")
    got = zs.ZarrArray(str(tmp_path / "A")).read(threads=2)
    assert np.array_equal(got[:2], a[:2]) and np.isnan(got[2:4]).all() and np.array_equal(got[4:], a[4:])
    # corrupt frames are errors, not garbage
    good = open(tmp_path / "A" / "0.0", "rb").read()
    with open(tmp_path / "A" / "0.0", "wb") as f:
        f.write(good[:-3])
    with pytest.raises(_lib.CngiError, match="blosc header size"):
        zs.ZarrArray(str(tmp_path / "A")).read(threads=1)
    with open(tmp_path / "A" / "0.0", "wb") as f:
        f.write(good[:2] + bytes([good[2] | 0x4]) + good[3:])                # bit shuffle
    with pytest.raises(_lib.CngiError, match="bit shuffle"):
        zs.ZarrArray(str(tmp_path / "A")).read(threads=1)
    zs.write_array(str(tmp_path / "R"), a, chunks=(2, 4))
    with open(tmp_path / "R" / "0.0", "ab") as f:
        f.write(b"x")
    with pytest.raises(_lib.CngiError, match="wrong size"):
        zs.ZarrArray(str(tmp_path / "R")).read(threads=1)
    # argument checks of the C entry point
    L = _lib.lib()
    job = (_lib.ZarrChunkJob * 1)()
    job[0].chunk_shape[0], job[0].extent[0] = 4, 5                           # box leaves its chunk
    out = np.zeros(8)
    import ctypes as C
    shape = (C.c_int64 * 1)(8)
    assert L.cngi_b200_zarr_read_chunks(job, 1, out.ctypes.data, shape, 1, 8, _lib.ZARR_RAW, out.ctypes.data, 1) == 1
    assert L.cngi_b200_zarr_read_chunks(job, 1, out.ctypes.data, shape, 1, 8, 7, out.ctypes.data, 1) == 1
    assert L.cngi_b200_zarr_read_chunks(None, 0, None, None, 1, 8, 0, None, 1) == 0


def test_host_chunk_walk_native_and_python_agree(tmp_path):
    d = _vis()
    store = rv.write_vis(str(tmp_path / "v.zarr"), d, chunks={"time": 4, "chan": 2})
    xds = rv.read_vis(store, partition="xds0").xds0
    a = list(xds.iter_host_chunks(time_chunk=5, workers=3, native=True))
    b = list(xds.iter_host_chunks(time_chunk=5, workers=3, native=False))
    assert [s for s, _ in a] == [s for s, _ in b]
    for (_, x), (_, y) in zip(a, b):
        assert set(x) == set(y) == {"DATA", "UVW", "WEIGHT", "FLAG", "FIELD_ID"}
        for k in x:
            assert x[k].dtype == y[k].dtype and np.array_equal(x[k], y[k]), k


@pytest.mark.parametrize("world_size", [2, 3])
def test_rank_shards_read_only_their_part_of_the_store(tmp_path, world_size, monkeypatch):
    """distributed.time_shard / channel_shard slice lazy zarr arrays like in-memory ones, so each rank of a multi-GPU
    job reads only the chunk files that intersect its shard (samples shard, nothing is exchanged
    before the grid reduce)."""
    from cngi_prototype_b200 import distributed as D
    d = _vis(shape=(12, 7, 6, 2))
    store = rv.write_vis(str(tmp_path / "v.zarr"), d, chunks={"time": 2, "chan": 2}, compressor=None)
    xds = rv.read_vis(store, partition="xds0").xds0
    lazy = {"vis": xds["DATA"], "uvw": xds["UVW"], "weight": xds["WEIGHT"], "freq_chan": xds["chan"]}
    mem = {"vis": d["DATA"], "uvw": d["UVW"], "weight": d["WEIGHT"], "freq_chan": d["chan"]}
    opened = []
    real = zs.ZarrArray.read_chunk

    def spy(self, idx):
        opened.append((os.path.basename(self.path), idx))
        return real(self, idx)

    monkeypatch.setattr(zs.ZarrArray, "read_chunk", spy)
    for rank in range(world_size):
        a, b = D.time_shard(lazy, rank, world_size), D.time_shard(mem, rank, world_size)
        for k in ("vis", "uvw", "weight"):
            assert np.array_equal(np.asarray(a[k]), b[k]), k
        lo, hi = D.shard_range(12, rank, world_size)
        mine = {i for n, i in opened if n == "DATA"}
        assert mine and all(lo // 2 <= i[0] <= (hi - 1) // 2 for i in mine)      # only this rank's time chunks
        opened.clear()
        a, b = D.channel_shard(lazy, rank, world_size), D.channel_shard(mem, rank, world_size)
        assert np.array_equal(np.asarray(a["vis"]), b["vis"]) and np.array_equal(np.asarray(a["freq_chan"]), b["freq_chan"])
        lo, hi = D.shard_range(6, rank, world_size)
        mine = {i for n, i in opened if n == "DATA"}
        assert mine and all(lo // 2 <= i[2] <= (hi - 1) // 2 for i in mine)      # only this rank's channel chunks
        opened.clear()
