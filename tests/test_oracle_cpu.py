"""CPU: pins the oracle against the golden vectors generated from the reference's own code
(tests/golden/make_golden.py) and against the hand-computed KATs of SURVEY.md section 8c."""
import json
import os

import numpy as np
import pytest

from oracle import oracle
from tests import inputs

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def test_pack_layout_matches_reference_goldens():
    for case in load("bits.json")["pack"]:
        rows = np.array(case["rows"], dtype=np.uint8)
        want = np.array([[int(x, 16) for x in r] for r in case["packed_hex"]], dtype=np.uint64)
        assert np.array_equal(oracle.py_pack_rows(rows), want), case["name"]


def test_k1_k2_literal():
    assert oracle.py_pack_rows(np.array([[1, 0], [0, 1], [1, 1]])).tolist() == [[0xA000000000000000, 0x6000000000000000]]
    rows = np.array([[1 if g in (0, 63, 64) else 0] for g in range(65)])
    assert oracle.py_pack_rows(rows).tolist() == [[0x8000000000000001], [0x8000000000000000]]


def test_minimum_uint_size_matches_reference():
    for c in load("bits.json")["minimum_uint_size"]:
        assert np.dtype(oracle.minimum_uint_size(int(c["value"]))).name == c["dtype"]


def test_oracle_matrix_uses_reference_bit_layout():
    """C oracle's packed matrix == reference packer applied to its own presence rows."""
    rng = np.random.default_rng(1)
    shared = [inputs.rand_seq(rng, 300)]
    genomes = [[(inputs.fasta(rng, shared=shared, max_len=120), 0)] for _ in range(70)]
    r = oracle.build(genomes, 9, 1, True)
    present = np.zeros((70, r.n_kmers), dtype=np.uint8)
    pos = {int(x): j for j, x in enumerate(r.kmers)}
    for g, files in enumerate(genomes):
        for x in oracle.py_genome_counts(files, 9):
            present[g, pos[x]] = 1
    assert np.array_equal(oracle.py_pack_rows(present), r.matrix)


def test_k4_known_answer():
    fa = b">r1\nGATTACAGGTNACCTGTAATC\n"
    k, c, nb, nw = oracle.genome_solid([(fa, 0)], 5)
    assert [int(x) for x in k] == [0x04F, 0x05B, 0x0A1, 0x1B8, 0x209, 0x284]
    assert c.tolist() == [2] * 6 and nb == 21 and nw == 12
    assert [oracle.py_kmer_string(int(x), 5) for x in k] == ["ACAGG", "ACCTG", "ATTAC", "CTGTA", "TAATC", "TTACA"]
    # the lexicographic-ACGT convention would have produced GATTA / TGTAA instead of TAATC / TTACA
    assert oracle.strand_neutral("TAATC") == "GATTA" and oracle.strand_neutral("TTACA") == "TGTAA"


@pytest.mark.parametrize("k", [1, 4, 11, 31, 32])
@pytest.mark.parametrize("kind", [0, 1])
def test_c_oracle_equals_python_restatement(k, kind):
    rng = np.random.default_rng(100 * k + kind)
    shared = [inputs.rand_seq(rng, 400)]
    genomes = []
    for g in range(5):
        if kind == 0:
            genomes.append([(inputs.fasta(rng, shared=shared, blank=True, crlf=(g == 1), final_nl=(g != 2),
                                          junk_prefix=(g == 3)), 0)])
        else:
            genomes.append([(inputs.fastq(rng, shared[0], n_reads=60, read_len=40, crlf=(g == 1), final_nl=(g != 2)), 1)])
    for m, keep in [(1, True), (1, False), (2, True)]:
        r = oracle.build(genomes, k, m, keep)
        pk, pm = oracle.py_build(genomes, k, m, keep)
        assert np.array_equal(r.kmers, pk) and np.array_equal(r.matrix, pm)


def test_threads_do_not_change_result():
    rng = np.random.default_rng(4)
    shared = [inputs.rand_seq(rng, 5000)]
    genomes = [[(inputs.fasta(rng, shared=shared, max_len=3000), 0)] for _ in range(9)]
    a = oracle.build(genomes, 15, 1, False, threads=1)
    b = oracle.build(genomes, 15, 1, False, threads=4)
    assert np.array_equal(a.kmers, b.kmers) and np.array_equal(a.matrix, b.matrix)


def test_tsv_grammar_is_what_from_tsv_expects():
    """create.py:121-137: k = len(first field of line 2); (size - header) % len(line 2) == 0."""
    rng = np.random.default_rng(2)
    genomes = [[(inputs.fasta(rng, max_len=200), 0)] for _ in range(4)]
    names = ["a", "bb", "c c", "d"]
    r = oracle.build(genomes, 12, 1, True, want_tsv_names=names)
    lines = r.tsv.split(b"\n")
    assert lines[0] == b"kmers\ta\tbb\tc c\td"
    assert len(lines[1].split(b"\t")[0]) == 12
    header = len(lines[0]) + 1
    assert (len(r.tsv) - header) % (len(lines[1]) + 1) == 0
    assert (len(r.tsv) - header) // (len(lines[1]) + 1) == r.n_kmers
    assert r.tsv == oracle.py_format_tsv(r.kmers, r.matrix, names, 12)


def test_synthetic_read_sets_and_min_abundance():
    """The C4 generator (grm_b200.synth.genome_reads_fastq) is deterministic, and on 30x reads with 0.5 % substitution
    errors min abundance 2 removes most of the k-mers (the error k-mers) while keeping the genome's own."""
    from grm_b200 import synth
    cfg = synth.SynthConfig(seed=synth.MASTER_SEED + 3).scaled(0.002)
    L = len(synth.genome_sequence(cfg, 0))
    fq = synth.genome_reads_fastq(cfg, 0, 30 * L // 150)
    assert fq == synth.genome_reads_fastq(cfg, 0, 30 * L // 150)
    assert fq.count(b"\n") == 4 * (30 * L // 150) and fq.startswith(b"@g0_r0\n")
    fa = synth.genome_fasta(cfg, 0)
    genome = oracle.build([[(fa, 0)]], 21, 1, True)
    all_k = oracle.build([[(fq, 1)]], 21, 1, True)
    solid = oracle.build([[(fq, 1)]], 21, 2, True)
    assert len(all_k.kmers) > 2 * len(solid.kmers)
    # contigs cut the genome, reads do not: the solid read k-mers cover (almost) every genome k-mer
    assert np.isin(genome.kmers, solid.kmers).mean() > 0.95
    assert np.isin(solid.kmers, all_k.kmers).all()
