"""CPU: hdf5min writer/reader -- byte-level checks against the HDF5 file format specification
(version-0 superblock, v1 object headers, symbol-table groups, v1 chunk B-trees, deflate) plus
round trips of every dtype / layout a .kover file uses."""
import struct
import zlib

import numpy as np
import pytest

from grm_b200 import hdf5min


def test_superblock_and_root_group_bytes(tmp_path):
    p = tmp_path / "a.h5"
    with hdf5min.H5Writer(str(p)) as h5:
        h5.attrs["uuid"] = "abc"
        h5.create_dataset("zeta", np.arange(3, dtype=np.uint8))
        h5.create_dataset("alpha", np.arange(5, dtype=np.uint64))
    b = p.read_bytes()
    assert b[:8] == b"\x89HDF\r\n\x1a\n"
    assert b[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])             # versions, size of offsets/lengths
    leaf_k, internal_k, flags = struct.unpack_from("<HHI", b, 16)
    assert (leaf_k, internal_k, flags) == (4, 16, 0)
    base, free, eof, driver = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and free == hdf5min.UNDEF and driver == hdf5min.UNDEF and eof == len(b)
    name_off, hdr, cache, _ = struct.unpack_from("<QQII", b, 56)
    assert name_off == 0 and cache == 1 and hdr % 8 == 0
    tree, heap = struct.unpack_from("<QQ", b, 80)
    assert b[tree:tree + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP"
    # object header v1: version 1, 16-byte prefix, messages 8-byte aligned
    ver, _, nmsg, refs, size = struct.unpack_from("<BBHII", b, hdr)
    assert ver == 1 and refs == 1 and size % 8 == 0 and nmsg == 2
    # symbol node entries are sorted by name (libhdf5 binary-searches them)
    r = hdf5min.H5Reader(str(p))
    assert list(r.datasets) == ["alpha", "zeta"]
    # group B-tree: one child, key[1] = heap offset of the largest name
    ntype, level, used = struct.unpack_from("<BBH", b, tree + 4)
    assert (ntype, level, used) == (0, 0, 1)
    heap_data = struct.unpack_from("<Q", b, heap + 24)[0]
    key1 = struct.unpack_from("<Q", b, tree + 24 + 16)[0]
    assert b[heap_data + key1:heap_data + key1 + 5] == b"zeta\0"
    # local heap free list: one block, H5HL_FREE_NULL terminated
    dsize, free_off, _ = struct.unpack_from("<QQQ", b, heap + 8)
    nxt, fsz = struct.unpack_from("<QQ", b, heap_data + free_off)
    assert nxt == 1 and free_off + fsz == dsize


@pytest.mark.parametrize("dtype", ["u1", "u2", "u4", "u8", "S1", "S31", "f8"])
def test_contiguous_round_trip(tmp_path, dtype):
    rng = np.random.default_rng(3)
    if dtype.startswith("S"):
        n = int(dtype[1:])
        a = np.array([bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)) for _ in range(17)], dtype=dtype)
    elif dtype == "f8":
        a = rng.random(9)
    else:
        a = rng.integers(0, np.iinfo(dtype).max, size=23, dtype=dtype, endpoint=True)
    p = tmp_path / "c.h5"
    with hdf5min.H5Writer(str(p)) as h5:
        h5.create_dataset("d", a, attrs={"description": "some text", "x": 1.5})
    d = hdf5min.H5Reader(str(p))["d"]
    assert d.dtype == a.dtype and d.shape == a.shape and np.array_equal(d.read(), a)
    assert d.attrs == {"description": "some text", "x": 1.5} and d.chunks is None


@pytest.mark.parametrize("gzip", [0, 1, 9])
@pytest.mark.parametrize("shape,chunks", [((3, 1000), (1, 256)), ((1, 10), (1, 10)), ((16, 70001), (1, 100000)),
                                           ((2, 129), (1, 1))])
def test_chunked_matrix_round_trip(tmp_path, gzip, shape, chunks):
    rng = np.random.default_rng(shape[1])
    a = rng.integers(0, 1 << 63, size=shape, dtype=np.uint64)
    p = tmp_path / "m.h5"
    with hdf5min.H5Writer(str(p), threads=3) as h5:
        h5.create_dataset("kmer_matrix", a, chunks=chunks, gzip=gzip)
    b = p.read_bytes()
    r = hdf5min.H5Reader(str(p))
    d = r["kmer_matrix"]
    eff = tuple(min(c, s) for c, s in zip(chunks, shape))
    assert d.chunks == eff and np.array_equal(d.read(), a)
    assert bool(d._filters) == (gzip > 0)
    if gzip:
        assert d._filters[0] == (1, (gzip,))                       # filter id 1 = deflate, one client value
    # every chunk record: size matches the stored blob; offsets are multiples of the chunk shape, sorted
    recs = r._walk_chunk_btree(d._layout["btree"], 2)
    n_chunks = -(-shape[0] // eff[0]) * -(-shape[1] // eff[1])
    assert len(recs) == n_chunks and [o for o, _, _ in recs] == sorted(o for o, _, _ in recs)
    for offs, addr, nbytes in recs:
        assert all(o % c == 0 for o, c in zip(offs, eff))
        raw = b[addr:addr + nbytes]
        if gzip:
            raw = zlib.decompress(raw)
        assert len(raw) == eff[0] * eff[1] * 8                     # edge chunks are stored full size
    if n_chunks > 64:                                              # more than 2K entries -> a two-level tree
        assert b[d._layout["btree"] + 5] >= 1


def test_empty_and_string_edge_cases(tmp_path):
    p = tmp_path / "e.h5"
    with hdf5min.H5Writer(str(p)) as h5:
        h5.attrs.update({"created": 123.25, "empty": "", "filter": "singleton"})
        h5.create_dataset("kmer_matrix", np.zeros((2, 0), dtype=np.uint64), chunks=(1, 1), gzip=4)
        h5.create_dataset("kmer_sequences", np.zeros(0, dtype="S31"), gzip=4)
        h5.create_dataset("phenotype_tags", np.zeros(0, dtype="S1"))
        h5.create_dataset("genome_identifiers", np.array(["a", "bbb"]), gzip=4)     # unicode -> S
    r = hdf5min.H5Reader(str(p))
    assert r.attrs["created"] == 123.25 and r.attrs["filter"] == "singleton" and r.attrs["empty"] == ""
    assert r["kmer_matrix"].shape == (2, 0) and r["kmer_matrix"].read().shape == (2, 0)
    assert r["kmer_sequences"].dtype == np.dtype("S31") and r["kmer_sequences"].read().size == 0
    assert r["genome_identifiers"].read().tolist() == [b"a", b"bbb"]


def test_root_group_capacity_is_enforced(tmp_path):
    h5 = hdf5min.H5Writer(str(tmp_path / "f.h5"))
    for i in range(9):
        h5.create_dataset(f"d{i}", np.arange(2, dtype=np.uint8))
    with pytest.raises(ValueError):
        h5.close()


def test_libhdf5_reads_what_hdf5min_writes(tmp_path):
    """When h5py (i.e. libhdf5, the library Kover's ds.py reads datasets with) is importable, it must read a
    .kover-shaped file written by hdf5min: attributes, fixed-length strings, a gzip-chunked matrix with an edge chunk.
    There is no h5py in the build image (the test is skipped there); it runs wherever the package is installed."""
    h5py = pytest.importorskip("h5py")
    rng = np.random.default_rng(5)
    mat = rng.integers(0, 1 << 63, size=(3, 250_001), dtype=np.uint64)
    seqs = np.array([b"ACGT" * 7 + b"ACG"] * 1000, dtype="S31")
    p = tmp_path / "d.kover"
    with hdf5min.H5Writer(str(p)) as h5:
        h5.attrs["created"] = "now"
        h5.attrs["compression"] = "gzip (level 4)"
        h5.create_dataset("genome_identifiers", np.array([b"562.1", b"562.22"], dtype="S6"))
        h5.create_dataset("kmer_sequences", seqs, chunks=(500,), gzip=4)
        h5.create_dataset("kmer_matrix", mat, chunks=(1, 100_000), gzip=4)
        h5.create_dataset("kmer_by_matrix_column", np.arange(250_001, dtype=np.uint32), chunks=(100_000,), gzip=4)
    with h5py.File(str(p), "r") as f:
        assert f.attrs["created"] in ("now", b"now")
        assert [x for x in f["genome_identifiers"][...]] == [b"562.1", b"562.22"]
        assert np.array_equal(f["kmer_sequences"][...], seqs)
        assert f["kmer_matrix"].chunks == (1, 100_000) and f["kmer_matrix"].compression == "gzip"
        assert np.array_equal(f["kmer_matrix"][...], mat)
        assert np.array_equal(f["kmer_by_matrix_column"][-3:], np.arange(249_998, 250_001))
