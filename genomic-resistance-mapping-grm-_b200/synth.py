"""Deterministic synthetic bacterial-scale genomes (no network, so no PATRIC downloads).

Counter-based generator: every base is a pure function of (seed, genome, position), so the
numpy code here and the CUDA kernel ``k_synth_fasta`` (csrc/grmkm_synth.cuh) materialise
identical FASTA bytes.  Spec: SURVEY.md section 8d, BASELINE.md section 4.

  core          core_len bases shared by all genomes, base = h(0,p) & 3
  shared sites  1 % of core positions; alt allele carried by genome g with frequency 2^-(1..7)
  private SNPs  1e-4 per base per genome (genome-unique k-mers -> singletons)
  accessory     n_islands x island_len pool, island present in g with probability 1/4
  contigs       n_contigs per genome, every second one reverse-complemented, 60-column FASTA
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

MASTER_SEED = 0x47524D31  # "GRM1"
MAGIC = 0x31304E59534D5247  # "GRMSYN01"
HEADER_WORDS = 16
_U = np.uint64
_C1 = _U(0x9E3779B97F4A7C15)
_C2 = _U(0xC2B2AE3D27D4EB4F)
_T_SITE = _U(184467440737095516)     # 2^64 / 100
_T_PRIV = _U(1844674407370955)       # 2^64 * 1e-4


@dataclass(frozen=True)
class SynthConfig:
    seed: int = MASTER_SEED
    core_len: int = 4_500_000
    island_len: int = 5_000
    n_islands: int = 400
    n_contigs: int = 50
    line_width: int = 60

    def scaled(self, factor: float) -> "SynthConfig":
        """Same structure at a smaller scale (tests)."""
        return SynthConfig(self.seed, max(200, int(self.core_len * factor)), max(50, int(self.island_len * factor)),
                           self.n_islands, self.n_contigs, self.line_width)


def splitmix64(x):
    with np.errstate(over="ignore"):
        x = np.asarray(x, dtype=_U) + _C1
        z = x
        z = (z ^ (z >> _U(30))) * _U(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> _U(27))) * _U(0x94D049BB133111EB)
        return z ^ (z >> _U(31))


def h(seed: int, a: int, b):
    with np.errstate(over="ignore"):
        return splitmix64(_U(seed) ^ (_U(a) * _C1) ^ (np.asarray(b, dtype=_U) * _C2))


def present_islands(cfg: SynthConfig, g: int) -> np.ndarray:
    i = np.arange(cfg.n_islands, dtype=_U)
    return np.nonzero((h(cfg.seed, 7, (_U(g) << _U(32)) + i) >> _U(62)) == 0)[0].astype(np.int64)


def contig_bounds(cfg: SynthConfig, g: int, genome_len: int) -> np.ndarray:
    C = cfg.n_contigs
    j = np.arange(1, C, dtype=_U)
    mod = max(1, genome_len // (4 * C))
    jitter = (h(cfg.seed, 9, (_U(g) << _U(32)) + j) % _U(mod)).astype(np.int64)
    b = np.empty(C + 1, dtype=np.int64)
    b[0] = 0
    b[C] = genome_len
    b[1:C] = (genome_len * np.arange(1, C, dtype=np.int64)) // C + jitter
    return b


def core_bases(cfg: SynthConfig, g: int, p: np.ndarray) -> np.ndarray:
    """bases (0..3 = A C G T) of genome g at core positions p."""
    p = np.asarray(p, dtype=_U)
    s = cfg.seed
    gp = (_U(g) << _U(32)) + p
    v = (h(s, 0, p) & _U(3)).astype(np.int64)
    site = h(s, 1, p) < _T_SITE
    e = _U(1) + h(s, 3, p) % _U(7)
    carry = (h(s, 4, gp) >> (_U(64) - e)) == 0
    alt = (v + 1 + (h(s, 2, p) % _U(3)).astype(np.int64)) & 3
    v = np.where(site & carry, alt, v)
    priv = h(s, 5, gp) < _T_PRIV
    alt2 = (v + 1 + (h(s, 8, gp) % _U(3)).astype(np.int64)) & 3
    return np.where(priv, alt2, v)


def genome_sequence(cfg: SynthConfig, g: int) -> np.ndarray:
    """linear genome (core followed by the present islands) as values 0..3."""
    parts = [core_bases(cfg, g, np.arange(cfg.core_len, dtype=_U))]
    q = np.arange(cfg.island_len, dtype=_U)
    for i in present_islands(cfg, g):
        parts.append((h(cfg.seed, 6, (_U(int(i)) << _U(32)) + q) & _U(3)).astype(np.int64))
    return np.concatenate(parts)


def contig_header(g: int, j: int) -> bytes:
    return f">g{g}_c{j}\n".encode()


def genome_fasta(cfg: SynthConfig, g: int) -> bytes:
    """FASTA text of genome g (numpy reference of k_synth_fasta)."""
    seq = genome_sequence(cfg, g)
    b = contig_bounds(cfg, g, len(seq))
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = []
    LW = cfg.line_width
    for j in range(cfg.n_contigs):
        c = seq[b[j]:b[j + 1]]
        if j & 1:
            c = 3 - c[::-1]
        L = len(c)
        nl = (L + LW - 1) // LW
        body = np.full(L + nl, ord("\n"), dtype=np.uint8)
        idx = np.arange(L)
        body[idx + idx // LW] = letters[c]
        out.append(contig_header(g, j))
        out.append(body.tobytes())
    return b"".join(out)


def genome_reads_fastq(cfg: SynthConfig, g: int, n_reads: int, read_len: int = 150, err: float = 0.005) -> bytes:
    """FASTQ read set of genome g: n_reads reads of read_len bases sampled from the genome's
    linear sequence, start and strand from the counter hash, substitution errors at rate err (mostly count-1 k-mers: what
    min-abundance 2 removes), 4-line records, constant quality 'I', header @g{g}_r{j}."""
    seq = genome_sequence(cfg, g)
    L = len(seq)
    read_len = min(read_len, L)
    j = np.arange(n_reads, dtype=_U)
    gj = (_U(g) << _U(32)) + j
    start = (h(cfg.seed, 10, gj) % _U(L - read_len + 1)).astype(np.int64)
    strand = (h(cfg.seed, 11, gj) & _U(1)).astype(bool)
    idx = start[:, None] + np.arange(read_len, dtype=np.int64)[None, :]
    bases = seq[idx]                                                   # [n_reads][read_len], 0..3 = A C G T
    bases = np.where(strand[:, None], 3 - bases[:, ::-1], bases)
    cell = gj[:, None] * _U(1024) + np.arange(read_len, dtype=_U)[None, :]
    hit = h(cfg.seed, 12, cell) < _U(int(err * 2.0 ** 64))
    sub = (bases + 1 + (h(cfg.seed, 13, cell) % _U(3)).astype(np.int64)) & 3
    bases = np.where(hit, sub, bases)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)[bases]            # uint8 [n_reads][read_len]
    qual = b"I" * read_len
    out = []
    for r in range(n_reads):
        out.append(b"@g%d_r%d\n" % (g, r))
        out.append(letters[r].tobytes())
        out.append(b"\n+\n" + qual + b"\n")
    return b"".join(out)


def _reads_header(g: int, n_reads: int) -> tuple[bytes, int]:
    return b"@g%d_r" % g, len(str(max(n_reads - 1, 0)))


def genome_reads_fastq_fixed(cfg: SynthConfig, g: int, n_reads: int, read_len: int = 150, err: float = 0.005) -> bytes:
    """The read set of genome_reads_fastq with the read number zero-padded, so that every record of a genome has the
    same width and the CUDA generator (k_synth_fastq) can place any byte directly: identical bytes on both sides."""
    seq = genome_sequence(cfg, g)
    L = len(seq)
    read_len = min(read_len, L)
    prefix, D = _reads_header(g, n_reads)
    pl = len(prefix)
    recw = pl + D + 1 + read_len + 3 + read_len + 1
    out = np.empty((n_reads, recw), dtype=np.uint8)
    j = np.arange(n_reads, dtype=_U)
    gj = (_U(g) << _U(32)) + j
    start = (h(cfg.seed, 10, gj) % _U(L - read_len + 1)).astype(np.int64)
    strand = (h(cfg.seed, 11, gj) & _U(1)).astype(bool)
    thr = _U(int(err * 2.0 ** 64))
    out[:, :pl] = np.frombuffer(prefix, dtype=np.uint8)
    v = np.arange(n_reads, dtype=np.int64)
    for d in range(D - 1, -1, -1):
        out[:, pl + d] = ord("0") + v % 10
        v //= 10
    out[:, pl + D] = 10
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    step = max(1, (1 << 22) // max(read_len, 1))
    t = np.arange(read_len, dtype=np.int64)
    for a in range(0, n_reads, step):
        b = min(n_reads, a + step)
        bases = seq[start[a:b, None] + t[None, :]]
        bases = np.where(strand[a:b, None], 3 - bases[:, ::-1], bases)
        cell = gj[a:b, None] * _U(1024) + t.astype(_U)[None, :]
        hit = h(cfg.seed, 12, cell) < thr
        sub = (bases + 1 + (h(cfg.seed, 13, cell) % _U(3)).astype(np.int64)) & 3
        out[a:b, pl + D + 1: pl + D + 1 + read_len] = letters[np.where(hit, sub, bases)]
    o = pl + D + 1 + read_len
    out[:, o] = 10
    out[:, o + 1] = ord("+")
    out[:, o + 2] = 10
    out[:, o + 3: o + 3 + read_len] = ord("I")
    out[:, recw - 1] = 10
    return out.tobytes()


def build_reads_layout(cfg: SynthConfig, genome_ids, n_reads: int, read_len: int = 150,
                       err: float = 0.005) -> tuple[np.ndarray, int, list[tuple[int, int]]]:
    """Layout table of grmkm_synth_fasta_device for read sets (k_synth_fastq): (u64 table, total bytes, [(offset, length)])."""
    NI = cfg.n_islands
    stride = 10 + NI
    genome_ids = list(genome_ids)
    G = len(genome_ids)
    lay = np.zeros(HEADER_WORDS + G * stride, dtype=_U)
    spans = []
    off = 0
    for n, g in enumerate(genome_ids):
        isl = present_islands(cfg, g)
        Lg = cfg.core_len + len(isl) * cfg.island_len
        rl = min(read_len, Lg)
        prefix, D = _reads_header(g, n_reads)
        if len(prefix) > 16:
            raise ValueError("read header prefix longer than 16 bytes")
        recw = len(prefix) + D + 1 + rl + 3 + rl + 1
        flen = n_reads * recw
        blk = lay[HEADER_WORDS + n * stride: HEADER_WORDS + (n + 1) * stride]
        blk[0:8] = [g, off, flen, len(isl), Lg, recw, len(prefix), D]
        blk[8:10] = np.frombuffer(prefix + b"\0" * (16 - len(prefix)), dtype="<u8")
        blk[10:10 + len(isl)] = isl.astype(_U)
        spans.append((off, flen))
        off += (flen + 15) & ~15
    lay[0:10] = [MAGIC, cfg.seed, G, cfg.core_len, cfg.island_len, cfg.n_contigs, cfg.line_width, off, NI, stride]
    lay[10:14] = [1, min(read_len, cfg.core_len), int(err * 2.0 ** 64), n_reads]
    return lay, off, spans


def build_layout(cfg: SynthConfig, genome_ids) -> tuple[np.ndarray, int, list[tuple[int, int]]]:
    """Layout table for grmkm_synth_fasta_device.  Returns (u64 table, total bytes, [(offset, length)])."""
    C, NI, LW = cfg.n_contigs, cfg.n_islands, cfg.line_width
    stride = 5 + (C + 1) + (C + 1) + C + 4 * C + NI
    genome_ids = list(genome_ids)
    G = len(genome_ids)
    lay = np.zeros(HEADER_WORDS + G * stride, dtype=_U)
    spans = []
    off = 0
    for n, g in enumerate(genome_ids):
        isl = present_islands(cfg, g)
        Lg = cfg.core_len + len(isl) * cfg.island_len
        b = contig_bounds(cfg, g, Lg)
        blk = lay[HEADER_WORDS + n * stride: HEADER_WORDS + (n + 1) * stride]
        coff = np.zeros(C + 1, dtype=np.int64)
        for j in range(C):
            hdr = contig_header(g, j)
            if len(hdr) > 32:
                raise ValueError("contig header longer than 32 bytes")
            L = int(b[j + 1] - b[j])
            coff[j + 1] = coff[j] + len(hdr) + L + (L + LW - 1) // LW
            blk[5 + 2 * (C + 1) + j] = len(hdr)
            padded = hdr + b"\0" * (32 - len(hdr))
            blk[5 + 2 * (C + 1) + C + 4 * j: 5 + 2 * (C + 1) + C + 4 * j + 4] = np.frombuffer(padded, dtype="<u8")
        flen = int(coff[C])
        blk[0], blk[1], blk[2], blk[3], blk[4] = g, off, flen, len(isl), Lg
        blk[5:5 + C + 1] = b.astype(_U)
        blk[5 + C + 1:5 + 2 * (C + 1)] = coff.astype(_U)
        base = 5 + 2 * (C + 1) + C + 4 * C
        blk[base:base + len(isl)] = isl.astype(_U)
        spans.append((off, flen))
        off += (flen + 15) & ~15
    total = off
    lay[0:10] = [MAGIC, cfg.seed, G, cfg.core_len, cfg.island_len, C, LW, total, NI, stride]
    return lay, total, spans


def n_bases_of(cfg: SynthConfig, genome_ids) -> int:
    return sum(cfg.core_len + len(present_islands(cfg, g)) * cfg.island_len for g in genome_ids)
