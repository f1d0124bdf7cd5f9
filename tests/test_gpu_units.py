"""GPU parity of the super-k-mer ("unit") path (csrc/grmkm_units.cuh): the oracle's matrix bit for bit (-m gpu)."""
import numpy as np
import pytest

from oracle import oracle
from tests import inputs
from tests.test_gpu_parity import check

pytestmark = pytest.mark.gpu

COMP = bytes.maketrans(b"ACGT", b"TGCA")


def revcomp(s: bytes) -> bytes:
    return s.translate(COMP)[::-1]


def fasta_of(records, width=60):
    out = []
    for i, s in enumerate(records):
        out.append(b">r%d\n" % i)
        out.extend(s[j:j + width] + b"\n" for j in range(0, len(s), width))
    return b"".join(out)


def shared_population(rng, G, core_len=40_000, snp=0.002):
    """G genomes = one core with a few SNPs each, cut into contigs at genome-specific places, every second
    contig reverse-complemented: what the unit path is built for."""
    core = bytearray(inputs.rand_seq(rng, core_len))
    genomes = []
    for g in range(G):
        s = bytearray(core)
        for p in np.nonzero(rng.random(core_len) < snp)[0]:
            s[p] = b"ACGT"[(b"ACGT".index(s[p]) + 1 + int(rng.integers(3))) & 3]
        cuts = sorted(int(x) for x in rng.integers(1, core_len, size=5))
        recs = [bytes(s[a:b]) for a, b in zip([0] + cuts, cuts + [core_len])]
        recs = [revcomp(r) if i & 1 else r for i, r in enumerate(recs)]
        genomes.append([fasta_of(recs, width=int(rng.integers(50, 90)))])
    return genomes


@pytest.mark.parametrize("k", [31, 21, 15, 32, 9])
def test_units_share_work_across_genomes(gpu, k):
    rng = np.random.default_rng(100 + k)
    genomes = shared_population(rng, 40)
    st = check(genomes, k, keep_singletons=False)
    assert st["n_units"] > 0 and st["n_unit_entries"] > 0
    if k >= 15:
        # 40 genomes of one species: the distinct units are a small fraction of all units, whatever the contig
        # layout and strand (unit boundaries depend on the sequence only)
        assert st["n_unit_entries"] * 8 < st["n_units"], st
        assert st["n_wide"] * 8 < st["n_windows"], st
    check(genomes, k, keep_singletons=True)


def test_exact_offsets_on_a_shared_population(gpu, monkeypatch):
    rng = np.random.default_rng(7)
    genomes = shared_population(rng, 12, core_len=20_000)
    genomes.append([inputs.fasta(rng, n_records=5, max_len=3000, p_n=0.05)])
    monkeypatch.setenv("GRMKM_EXACT_OFFSETS", "1")
    for k in (31, 13, 5):
        st = check(genomes, k, keep_singletons=True)
        assert st["n_units"] > 0
        check(genomes, k, keep_singletons=False)


def test_low_complexity_and_ties(gpu):
    """Homopolymers and tandem repeats make equal minimizers enter and leave the window all the time: a run
    must still end within lmax k-mers, and every window must be in exactly one unit."""
    rng = np.random.default_rng(3)
    recs = [b"A" * 5000, b"ACACACACAC" * 300, b"T" * 777 + inputs.rand_seq(rng, 500) + b"G" * 900,
            (b"ACGGT" * 7 + inputs.rand_seq(rng, 11)) * 40, inputs.rand_seq(rng, 3000), b"C" * 31, b"C" * 30,
            b"GATTACA" * 500]
    genomes = [[fasta_of(recs)], [fasta_of([revcomp(r) for r in recs[::-1]], width=71)], [fasta_of(recs[2:5], width=33)]]
    for k in (31, 32, 22, 21, 15, 7, 2):
        check(genomes, k, keep_singletons=True)
    check(genomes, 31, keep_singletons=False)


def test_many_units_per_group_and_rounds(gpu):
    """Small k: nearly every k-mer is its own unit, so a thread has far more units than it holds in registers per
    round; N runs cut records into fragments shorter than k."""
    rng = np.random.default_rng(17)
    genomes = [[inputs.fasta(rng, n_records=6, min_len=2000, max_len=30_000, p_n=0.02, width=int(rng.integers(20, 200)))]
               for _ in range(9)]
    for k in (1, 3, 6, 8, 11):
        check(genomes, k, keep_singletons=bool(k & 1))


def test_dedupe_flushes_and_few_buckets(gpu, monkeypatch):
    """One unit bucket and more distinct units than the shared-memory table holds: the table is flushed several
    times, duplicates across flushes are merged by the column aggregate."""
    monkeypatch.setenv("GRMKM_UNIT_BUCKETS", "1")
    rng = np.random.default_rng(23)
    genomes = [[inputs.fasta(rng, n_records=2, min_len=150_000, max_len=200_000, p_n=0.0, p_iupac=0.0)] for _ in range(3)]
    genomes += [[g[0]] for g in genomes[:2]]          # two exact copies: shared with rows 0 and 1
    st = check(genomes, 31, keep_singletons=False)
    assert st["n_unit_buckets"] == 1 and st["n_unit_entries"] > 3 * 6144
    monkeypatch.setenv("GRMKM_UNIT_BUCKETS", "7")
    check(genomes, 21, keep_singletons=True)


def test_unit_region_overflow_falls_back_to_exact(gpu):
    """Thousands of copies of one short record: all their units are identical, so they hash to the same few
    buckets and overflow the over-provisioned regions -> the build is redone with exact offsets."""
    rng = np.random.default_rng(29)
    rec = inputs.rand_seq(rng, 90)
    genomes = [[fasta_of([rec] * 6000)], [fasta_of([revcomp(rec)] * 3000 + [inputs.rand_seq(rng, 5000)])]]
    st = check(genomes, 31, keep_singletons=True)
    assert st["n_region_overflows"] >= 1
    check(genomes, 31, keep_singletons=False)


def test_units_across_word_blocks(gpu):
    """More than 64 genomes: a unit's presence is kept per 64-genome block."""
    rng = np.random.default_rng(31)
    genomes = shared_population(rng, 150, core_len=6000, snp=0.004)
    check(genomes, 31, keep_singletons=False)
    check(genomes, 17, keep_singletons=True)


def test_fastq_without_abundance_filter_uses_units(gpu):
    rng = np.random.default_rng(37)
    src = [inputs.rand_seq(rng, 4000) for _ in range(2)]
    genomes = [[inputs.fastq(rng, src[g % 2], n_reads=600, read_len=80)] for g in range(5)]
    st = check(genomes, 21, min_abundance=1, keep_singletons=True, kind=1)
    assert st["n_units"] > 0
    st = check(genomes, 21, min_abundance=2, keep_singletons=True, kind=1)
    assert st["n_units"] > 0 and st["n_rounds"] == 2              # abundance builds use the units too, in rounds of 4 rows


def test_expansion_regions_too_small_fall_back_to_exact_offsets(gpu, monkeypatch):
    """The wide records go to over-provisioned bucket regions sized from an estimate (no count pass, no host round
    trip); an estimate that is far too small sets the overflow flag and the round is repeated with exact offsets."""
    rng = np.random.default_rng(41)
    genomes = shared_population(rng, 70, core_len=20000, snp=0.003)
    monkeypatch.setenv("GRMKM_WIDE_EST", "64")
    st = check(genomes, 31, keep_singletons=True)
    assert st["n_region_overflows"] >= 1 and st["n_wide"] > 20000
    monkeypatch.delenv("GRMKM_WIDE_EST")
    check(genomes, 31, keep_singletons=False)
    monkeypatch.setenv("GRMKM_WIDE_EXACT", "1")                      # the count-pass path by itself
    check(genomes, 21, keep_singletons=True)


def test_entry_list_too_small_is_retried(gpu, monkeypatch):
    """The dedupe's entry list is sized from an estimate as well; the build is repeated with the size it asked for,
    with the regions (first) and with exact offsets (second)."""
    rng = np.random.default_rng(43)
    genomes = shared_population(rng, 9, core_len=30000, snp=0.01)
    monkeypatch.setenv("GRMKM_WU_CAP", "100")
    st = check(genomes, 31, keep_singletons=True)
    assert st["n_unit_entries"] > 100
    monkeypatch.setenv("GRMKM_WIDE_EXACT", "1")
    check(genomes, 31, keep_singletons=False)


def test_more_buckets_than_the_expansion_sorts_become_key_sub_ranges(gpu, monkeypatch):
    """More distinct k-mers than 2^11 shared-memory tables hold: the unit path keeps 2^11 hash buckets and aggregates each
    as 2^s key sub-ranges (virtual buckets) instead of dropping to the per-record scatter."""
    rng = np.random.default_rng(47)
    genomes = shared_population(rng, 12, core_len=50000, snp=0.01)
    st = check(genomes, 31, keep_singletons=True, bucket_bits=13)          # 2^11 buckets x 2^2 sub-ranges
    assert st["n_units"] > 0 and st["n_buckets"] == 2048
    st = check(genomes, 21, keep_singletons=False, bucket_bits=15)
    assert st["n_units"] > 0 and st["n_buckets"] == 2048
    monkeypatch.setenv("GRMKM_SUB_BITS", "1")
    st = check(genomes, 31, keep_singletons=False)
    assert st["n_units"] > 0
    genomes = shared_population(rng, 130, core_len=4000, snp=0.01)         # three word-rows, two genome groups
    check(genomes, 31, keep_singletons=False)
