"""GPU parity of the kernels that run over the finished matrix (checksum, the learner's masked popcount row sums,
the Gram matrix, from_tsv's bit packer) and of the device read-set generator (-m gpu)."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle
from tests import inputs

pytestmark = pytest.mark.gpu


def built(genomes, k=21, keep=True):
    from grm_b200.builder import KmerMatrixBuilder
    b = KmerMatrixBuilder(k=k, keep_singletons=keep)
    b.set_genome_count(len(genomes))
    for row, fa in enumerate(genomes):
        b.add_genome_bytes(row, fa)
    b.build()
    return b


@pytest.mark.parametrize("G", [3, 64, 150])
def test_checksum_sum_rows_gram(gpu, G):
    rng = np.random.default_rng(G)
    shared = [inputs.rand_seq(rng, 3000), inputs.rand_seq(rng, 800)]
    genomes = [inputs.fasta(rng, n_records=2, max_len=600, shared=shared) for _ in range(G)]
    with built(genomes) as b:
        km, mat = b.kmers(), b.matrix()
        ref = oracle.build([[(g, 0)] for g in genomes], 21, 1, True)
        assert np.array_equal(km, ref.kmers) and np.array_equal(mat, ref.matrix)
        assert b.checksum() == oracle.checksum(ref.kmers, ref.matrix)
        # a permutation of the columns has the same digest, a flipped bit does not
        perm = rng.permutation(len(km))
        assert oracle.checksum(km[perm], mat[:, perm]) == b.checksum()
        bad = mat.copy(); bad[0, 0] ^= np.uint64(1) << np.uint64(63)
        assert oracle.checksum(km, bad) != b.checksum()
        # sum_rows: all rows, a random subset, one row, none
        U = len(km)
        for rows in (list(range(G)), sorted(rng.choice(G, size=max(1, G // 3), replace=False).tolist()), [G - 1], []):
            want = oracle.sum_rows(ref.matrix, rows, G)
            got = b.sum_rows(oracle.row_mask(rows, G))
            assert np.array_equal(got, want[:U].astype(np.uint32))
        assert np.array_equal(b.gram(), oracle.gram(ref.matrix, G))


def test_rule_classifications_mirror(gpu):
    """The reference-facing class: sum_rows returns presence and absence halves with the reference's dtype rule."""
    from grm_b200.learning import KmerRuleClassifications
    rng = np.random.default_rng(4)
    shared = [inputs.rand_seq(rng, 2500)]
    G = 70
    genomes = [inputs.fasta(rng, n_records=2, max_len=500, shared=shared) for _ in range(G)]
    with built(genomes) as b:
        ref = oracle.build([[(g, 0)] for g in genomes], 21, 1, True)
        rc = KmerRuleClassifications(b, G)
        assert rc.shape == (G, 2 * ref.n_kmers)
        rows = [0, 3, 64, 69]
        got = rc.sum_rows(rows)
        want = oracle.sum_rows(ref.matrix, rows, G)
        assert got.dtype == want.dtype and np.array_equal(got, want)
        cols = rc.get_columns([0, 5, ref.n_kmers + 5])
        assert cols.shape == (G, 3)
        assert np.array_equal(cols[:, 1] + cols[:, 2], np.ones(G, dtype=cols.dtype))       # a rule and its absence twin
        rc.remove_rows([1, 2])                          # rows are then counted among the remaining ones
        assert rc.shape == (G - 2, 2 * ref.n_kmers)
        got2 = rc.sum_rows([0, 1])                      # = original rows 0 and 3
        assert np.array_equal(got2, oracle.sum_rows(ref.matrix, [0, 3], G).astype(got2.dtype))


@pytest.mark.parametrize("G,U", [(3, 50), (64, 1000), (100, 40_000), (130, 777)])
def test_tsv_pack(gpu, G, U):
    """TSV text -> matrix words, with a column selection (the genomes kept after _parse_metadata, in label order)."""
    from grm_b200.builder import KmerMatrixBuilder
    rng = np.random.default_rng(G + U)
    k = 31
    cells = (rng.random((U, G)) < 0.4).astype(np.uint8)
    roww = k + 2 * G + 1
    body = np.full((U, roww), ord("\t"), dtype=np.uint8)
    body[:, :k] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(U, k))
    body[:, k + 1:roww - 1:2] = cells + ord("0")
    body[:, roww - 1] = ord("\n")
    sel = rng.permutation(G)[: max(1, G - 2)]
    with KmerMatrixBuilder(k=k) as b:
        got = b.tsv_pack(body, roww, k, G, sel)
        want = oracle.py_pack_rows(cells[:, sel].T.astype(np.uint64))
        assert got.shape == want.shape and np.array_equal(got, want)
        body[U // 2, k + 1 + 2 * int(sel[0])] = ord("2")
        from grm_b200.native import GrmkmError
        with pytest.raises(GrmkmError):
            b.tsv_pack(body, roww, k, G, sel)


def test_from_tsv_on_the_gpu_writes_the_same_file(gpu, tmp_path):
    """from_tsv through the GPU packer equals the host packer bit for bit (the file is compared dataset by dataset)."""
    from grm_b200 import create, hdf5min
    rng = np.random.default_rng(12)
    shared = [inputs.rand_seq(rng, 1500)]
    G = 9
    genomes = [inputs.fasta(rng, n_records=2, max_len=400, shared=shared) for _ in range(G)]
    names = [f"g{i}" for i in range(G)]
    with built(genomes, k=31) as b:
        tsv = b.tsv(names)
    p = tmp_path / "KmerMatrix.tsv"
    tsv.tofile(p)
    md = tmp_path / "md.tsv"
    md.write_text("".join(f"g{i}\t{i % 2}\n" for i in range(G) if i != 4))
    create.from_tsv(str(p), str(tmp_path / "gpu.kover"), "pheno", str(md), 4)
    create.from_tsv(str(p), str(tmp_path / "host.kover"), "pheno", str(md), 4, use_gpu=False)
    a, c = hdf5min.H5Reader(str(tmp_path / "gpu.kover")), hdf5min.H5Reader(str(tmp_path / "host.kover"))
    for name in ("kmer_matrix", "kmer_sequences", "genome_identifiers", "phenotype", "kmer_by_matrix_column"):
        assert np.array_equal(a[name].read(), c[name].read()), name


def test_synth_reads_device_matches_numpy(gpu):
    import torch
    from grm_b200 import synth
    from grm_b200.builder import KmerMatrixBuilder
    cfg = synth.SynthConfig(seed=synth.MASTER_SEED + 3).scaled(0.004)
    ids, n_reads = [0, 7, 123], 1234
    lay, total, spans = synth.build_reads_layout(cfg, ids, n_reads)
    with KmerMatrixBuilder(k=31, input_kind=1) as b:
        buf = torch.zeros(total, dtype=torch.uint8, device="cuda")
        b._check(b._lib.grmkm_synth_fasta_device(b._ctx, C.c_void_p(lay.ctypes.data), lay.nbytes, C.c_void_p(buf.data_ptr()), total))
        host = buf.cpu().numpy()
    for g, (off, n) in zip(ids, spans):
        want = synth.genome_reads_fastq_fixed(cfg, g, n_reads)
        assert len(want) == n
        assert host[off:off + n].tobytes() == want
    # the fixed-width read set is the variable-width one with padded read numbers
    a = synth.genome_reads_fastq_fixed(cfg, 7, 50).split(b"\n")
    v = synth.genome_reads_fastq(cfg, 7, 50).split(b"\n")
    assert a[1::4] == v[1::4] and a[3::4] == v[3::4]
