"""GPU parity of abundance builds (-abundance-min > 1, kmer_count.py:44-53: rounds of genome rows with per-genome
counters, then one presence merge) and of the pooled count table (dsk mode, src/app.py:1356-1416) -- bit-exact
against the oracle (-m gpu)."""
import numpy as np
import pytest

from oracle import oracle
from tests import inputs
from tests.test_gpu_parity import check, gpu_build

pytestmark = pytest.mark.gpu


def read_sets(rng, G, src_len=6000, n_reads=400, read_len=80, err=0.01, n_src=2):
    srcs = [inputs.rand_seq(rng, src_len) for _ in range(n_src)]
    return [[inputs.fastq(rng, srcs[int(rng.integers(n_src))], n_reads=n_reads, read_len=read_len, err=err)] for _ in range(G)]


@pytest.mark.parametrize("min_ab", [2, 3, 7])
@pytest.mark.parametrize("G", [1, 5, 70])
def test_rounds_small(gpu, G, min_ab):
    """1, 2 and 18 rounds (4 rows each, never across a 64-row word); rows 64.. land in matrix word 1."""
    rng = np.random.default_rng(10 * G + min_ab)
    genomes = read_sets(rng, G, n_reads=250 if G > 8 else 600)
    st = check(genomes, 21, min_abundance=min_ab, keep_singletons=True, kind=1)
    assert st["n_rounds"] == sum((min(64, G - w) + 3) // 4 for w in range(0, G, 64))
    check(genomes, 21, min_abundance=min_ab, keep_singletons=False, kind=1)


@pytest.mark.parametrize("rows", [1, 3, 16])
def test_round_width_and_empty_rows(gpu, monkeypatch, rows):
    """GRMKM_ROUND_ROWS changes how many genome rows share a round (counter planes per table slot); rows without
    input, rows with several files and a row whose reads are all too short yield no solid k-mer and still get a row."""
    monkeypatch.setenv("GRMKM_ROUND_ROWS", str(rows))
    rng = np.random.default_rng(77 + rows)
    genomes = read_sets(rng, 11)
    genomes[2] = []                                                          # no file at all
    genomes[5] = [genomes[5][0], inputs.fastq(rng, inputs.rand_seq(rng, 3000), n_reads=300, read_len=60)]
    genomes[7] = [inputs.fastq(rng, inputs.rand_seq(rng, 500), n_reads=50, read_len=20)]     # shorter than k
    for keep in (True, False):
        check(genomes, 31, min_abundance=2, keep_singletons=keep, kind=1)


def test_abundance_on_contigs_and_small_k(gpu):
    """The C ABI allows an abundance filter on FASTA too (repeats inside a genome), and k below the minimizer length."""
    rng = np.random.default_rng(5)
    rep = inputs.rand_seq(rng, 900)
    genomes = []
    for _ in range(9):
        body = b"".join(rep if rng.random() < 0.5 else inputs.rand_seq(rng, 700) for _ in range(8))
        genomes.append([b">c\n" + body + b"\n"])
    for k in (31, 9, 4):
        check(genomes, k, min_abundance=2, keep_singletons=True)
        check(genomes, k, min_abundance=3, keep_singletons=False)


def test_round_output_grows_and_table_splits(gpu, monkeypatch):
    """A tiny first guess for the rounds' record buffer (it doubles when a round does not fit) and few buckets (the
    counter tables overflow and split into key sub-ranges)."""
    rng = np.random.default_rng(6)
    genomes = read_sets(rng, 9, src_len=40_000, n_reads=4000, read_len=100, err=0.02)
    st = check(genomes, 25, min_abundance=2, keep_singletons=True, kind=1, bucket_bits=4)
    assert st["n_splits"] > 0
    assert st["n_solid_records"] > 0


def test_rebuild_after_abundance_build(gpu):
    """One context: abundance build, then another one with other inputs (buffers and hints are reused)."""
    from grm_b200.builder import KmerMatrixBuilder
    rng = np.random.default_rng(8)
    with KmerMatrixBuilder(k=21, min_abundance=2, keep_singletons=True, input_kind=1) as b:
        for rep in range(3):
            genomes = read_sets(rng, 6 + rep, n_reads=300 + 200 * rep)
            b.reset()
            b.set_genome_count(len(genomes))
            for row, files in enumerate(genomes):
                for f in files:
                    b.add_genome_bytes(row, f)
            b.build()
            ref = oracle.build([[(f, 1) for f in files] for files in genomes], 21, 2, True)
            assert np.array_equal(b.kmers(), ref.kmers) and np.array_equal(b.matrix(), ref.matrix)
            assert b.stats["n_windows"] == ref.n_windows and b.stats["n_bases"] == ref.n_bases


@pytest.mark.parametrize("kind,min_ab", [(0, 1), (0, 2), (1, 2), (1, 4)])
def test_pooled_count_table(gpu, kind, min_ab):
    """dsk mode: every file pooled, (k-mer, abundance) for the k-mers at or above the threshold -- the oracle's
    per-genome solid list of ONE genome made of all the files."""
    from grm_b200 import native
    from grm_b200.builder import KmerMatrixBuilder
    rng = np.random.default_rng(50 + kind + min_ab)
    shared = [inputs.rand_seq(rng, 5000)]
    if kind == 0:
        files = [inputs.fasta(rng, n_records=3, max_len=2500, shared=shared) for _ in range(7)]
    else:
        files = [inputs.fastq(rng, shared[0], n_reads=500, read_len=70) for _ in range(5)]
    for k in (31, 11):
        with KmerMatrixBuilder(k=k, min_abundance=min_ab, input_kind=kind, flags=native.FLAG_COUNTS) as b:
            for i, f in enumerate(files):
                b.add_genome_bytes(i, f)                    # the row is ignored: everything is pooled
            b.build()
            km, cnt, st = b.kmers(), b.matrix(), b.stats
        rk, rc, nb, nw = oracle.genome_solid([(f, kind) for f in files], k, min_ab)
        assert st["n_genomes"] == 1 and cnt.shape == (1, len(km))
        assert st["n_bases"] == nb and st["n_windows"] == nw
        perm = np.argsort(km)
        assert np.array_equal(km[perm], rk)
        assert np.array_equal(cnt[0][perm], rc.astype(np.uint64))
        assert np.array_equal(km, rk[oracle.column_order(rk)])      # columns in ascending hash order, like every build


def test_c4_full_size_read_sets(gpu):
    """BASELINE.json configs[3] at full per-genome size: 8 read sets of 30x 150 bp reads over the 5 Mbp synthetic genomes
    (1 M reads = 310 MB of FASTQ each, synthesised on the device), min abundance 2 -- against the oracle."""
    import ctypes as C
    import torch
    from grm_b200 import synth
    from grm_b200.builder import KmerMatrixBuilder
    cfg = synth.SynthConfig(seed=synth.MASTER_SEED + 3)
    G, n_reads = 8, 1_000_000
    with KmerMatrixBuilder(k=31, min_abundance=2, keep_singletons=True, input_kind=1) as b:
        lay, total, spans = synth.build_reads_layout(cfg, range(G), n_reads)
        buf = torch.empty(total, dtype=torch.uint8, device="cuda")
        b._check(b._lib.grmkm_synth_fasta_device(b._ctx, C.c_void_p(lay.ctypes.data), lay.nbytes, C.c_void_p(buf.data_ptr()), total))
        b.set_genome_count(G)
        for row, (off, n) in enumerate(spans):
            b.add_genome_device(row, buf.data_ptr() + off, n)
        b.build()
        st = b.stats
        host = buf.cpu().numpy()
        ref = oracle.build([[(host[o:o + n].tobytes(), 1)] for o, n in spans], 31, 2, True)
        assert st["n_bases"] == ref.n_bases == G * n_reads * 150
        assert st["n_windows"] == ref.n_windows
        assert np.array_equal(b.kmers(), ref.kmers) and np.array_equal(b.matrix(), ref.matrix)
        assert b.checksum() == oracle.checksum(ref.kmers, ref.matrix)
