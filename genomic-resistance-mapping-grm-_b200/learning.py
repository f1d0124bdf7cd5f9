"""The learner's hot loop on the GPU-resident matrix -- Python-3 stand-in for
``KmerRuleClassifications`` (bin/kover/core/kover/learning/common/rules.py:96-267) with the same methods,
argument meaning and result types.

The reference streams the packed ``kmer_matrix`` from HDF5 block by block and runs a Cython popcount over it
(popcount.pyx:76-95) for every call of ``sum_rows`` (split risks, split.py:176-188; SCM utilities, scm.py:238-288).
Here the matrix stays where the build left it (HBM) and one kernel does the masked popcount row reduction
(``grmkm_sum_rows``: 8 x ceil(G/64) bytes per k-mer, HBM-bound).  Rule j < U is "k-mer j present", rule U + j
its absence twin (rules.py:146-148, 264).
"""
from __future__ import annotations

import numpy as np

from .create import _minimum_uint_size


def _unpack_binary_bytes_from_ints(a: np.ndarray) -> np.ndarray:
    """(W x n) uint64 -> (64 W x n) 0/1 bytes, genome g = bit 63-(g%64) of word g//64 (kover/utils.py:159-187)."""
    a = np.ascontiguousarray(a, dtype=np.uint64)
    W, n = a.shape
    by = a.astype(">u8").view(np.uint8).reshape(W, n, 8)
    bits = np.unpackbits(by, axis=2)                       # (W, n, 64): bit 63 first
    return np.ascontiguousarray(bits.transpose(0, 2, 1)).reshape(W * 64, n)


class KmerRuleClassifications:
    """Rule classifications of a built matrix.  ``builder`` is a KmerMatrixBuilder (or DistributedBuilder rank)
    whose result is resident; ``n_rows`` the number of genomes (examples)."""

    def __init__(self, builder, n_rows: int):
        self.builder = builder
        self.dataset_initial_n_rows = int(n_rows)
        self.dataset_n_rows = int(n_rows)
        self.dataset_removed_rows: list[int] = []
        self.dataset_removed_rows_mask = np.zeros(self.dataset_initial_n_rows, dtype=bool)
        U, W, _ = builder.dims
        self._U, self._W = int(U), int(W)
        self._matrix = None

    # -- row bookkeeping (rules.py:171-199): rows are numbered among the rows not removed yet ------------
    def _absolute_rows(self, rows) -> list[int]:
        alive = np.nonzero(~self.dataset_removed_rows_mask)[0]
        return [int(alive[int(r)]) for r in rows]

    def remove_rows(self, rows):
        removed = self._absolute_rows(rows)
        if removed:
            self.dataset_removed_rows = sorted(set(self.dataset_removed_rows + removed))
            self.dataset_removed_rows_mask = np.zeros(self.dataset_initial_n_rows, dtype=bool)
            self.dataset_removed_rows_mask[self.dataset_removed_rows] = True
            self.dataset_n_rows = self.dataset_initial_n_rows - len(self.dataset_removed_rows)

    @property
    def shape(self):
        return self.dataset_n_rows, self._U * 2

    # -- sum_rows (rules.py:201-267) -----------------------------------------------------------------------
    def sum_rows(self, rows):
        """Number of the given rows (no duplicates) each rule classifies as positive: [presence rules | absence rules]."""
        rows = np.asarray(rows)
        result = np.zeros(self._U * 2, dtype=_minimum_uint_size(rows.shape[0]))
        mask = np.zeros(self._W, dtype=np.uint64)
        for idx in self._absolute_rows(np.sort(rows)):
            mask[idx // 64] |= np.uint64(1) << np.uint64(64 - (idx - 64 * (idx // 64)) - 1)      # build_row_mask, rules.py:209-222
        sums = self.builder.sum_rows(mask)
        result[: self._U] = sums
        result[self._U:] = len(rows) - sums.astype(np.int64)
        return result

    # -- get_columns (rules.py:135-171) --------------------------------------------------------------------
    def get_columns(self, columns):
        columns_is_int = False
        if hasattr(columns, "__index__"):
            columns = [columns.__index__()]
            columns_is_int = True
        columns = [int(c) for c in columns]
        invert = np.array([c >= self._U for c in columns])
        cols = np.array([c % self._U if c >= self._U else c for c in columns], dtype=np.int64)
        if self._matrix is None:
            self._matrix = self.builder.matrix()
        row_mask = np.ones(self._W * 64, dtype=bool)
        row_mask[self.dataset_initial_n_rows:] = False
        row_mask[self.dataset_removed_rows] = False
        result = _unpack_binary_bytes_from_ints(self._matrix[:, cols])[row_mask]
        result[:, invert] = 1 - result[:, invert]
        return result.reshape(-1) if columns_is_int else result
