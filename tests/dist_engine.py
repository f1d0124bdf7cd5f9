"""Oracle-backed stand-in for CudaEngine so the N>1 host logic runs on CPU (gloo) -- test only."""
import numpy as np
import torch

from oracle import oracle

M64 = (1 << 64) - 1


MUL = 0x9E3779B97F4A7C15            # the product's multiplicative hash (grmkm_device.cuh: khash / kunhash)
INV = pow(MUL, -1, 1 << 64)


def fmix64(h):
    return (h * MUL) & M64


def unfmix64(h):
    return (h * INV) & M64


class OracleEngine:
    def __init__(self, k, min_abundance, keep_singletons, kind=0):
        self.k, self.m, self.keep, self.kind = k, min_abundance, keep_singletons, kind
        self.files = {}
        self.bits = 6
        self.launches = 0

    # builder-like input surface
    def reset(self):
        self.files = {}

    def set_genome_count(self, n):
        self.n_local = n

    def add_genome_bytes(self, row, data):
        self.files.setdefault(row, []).append((bytes(data), self.kind))

    def close(self):
        pass

    # engine surface
    def plan_bucket_bits(self):
        return 6 + (len(self.files) % 3)      # ranks disagree on purpose: the driver must take the max

    def set_bucket_bits(self, bits):
        self.bits = bits

    def build_partial(self, world, wl):
        cols = {}
        for row, files in self.files.items():
            kmers, _, _, _ = oracle.genome_solid(files, self.k, self.m)
            for x in kmers.tolist():
                w = cols.setdefault(fmix64(x), [0] * wl)
                w[row >> 6] |= 1 << (63 - (row & 63))
        B = 1 << self.bits
        per_owner = [[] for _ in range(world)]
        for h, w in cols.items():
            owner = ((h >> (64 - self.bits)) * world) // B
            per_owner[owner].append([h] + w)
        counts = [len(p) for p in per_owner]
        flat = [v for p in per_owner for rec in p for v in rec]
        send = torch.from_numpy(np.array(flat, dtype=np.uint64).view(np.int64)) if flat else torch.empty(0, dtype=torch.int64)
        return counts, send

    def merge(self, recv, world, rank, src_counts, src_words, n_genomes):
        a = recv.numpy().view(np.uint64)
        W = sum(src_words)
        cols, pos, woff = {}, 0, 0
        for s in range(world):
            width = 1 + src_words[s]
            for i in range(src_counts[s]):
                rec = a[pos + i * width: pos + (i + 1) * width].tolist()
                w = cols.setdefault(rec[0], [0] * W)
                for j in range(src_words[s]):
                    w[woff + j] |= rec[1 + j]
            pos += src_counts[s] * width
            woff += src_words[s]
        keep = sorted((h, w) for h, w in cols.items() if self.keep or sum(bin(x).count("1") for x in w) >= 2)
        keep = [(unfmix64(h), w) for h, w in keep]           # ascending hash order, k-mers out
        self._k = np.array([k for k, _ in keep], dtype=np.uint64)
        self._m = np.array([w for _, w in keep], dtype=np.uint64).reshape(len(keep), W).T.copy()

    def result(self):
        return self._k, self._m
