"""Host-side Ray-Surveyor-format TSV writer (used when the columns were gathered from several GPUs;
on one GPU the text is formatted on the device by grmkm_format_tsv)."""
from __future__ import annotations

import numpy as np

_LETTERS = np.frombuffer(b"ACTG", dtype=np.uint8)     # GATB code order A0 C1 T2 G3


def kmer_strings(kmers: np.ndarray, k: int) -> np.ndarray:
    kmers = np.asarray(kmers, dtype=np.uint64)
    shifts = (2 * (k - 1 - np.arange(k))).astype(np.uint64)
    codes = (kmers[:, None] >> shifts[None, :]) & np.uint64(3)
    return np.ascontiguousarray(_LETTERS[codes.astype(np.int64)]).view(f"S{k}").reshape(-1)


def format_tsv(kmers: np.ndarray, matrix: np.ndarray, names, k: int) -> np.ndarray:
    G, U = len(names), len(kmers)
    header = ("kmers" + "".join("\t" + n for n in names) + "\n").encode()
    roww = k + 2 * G + 1
    body = np.empty((U, roww), dtype=np.uint8)
    if U:
        body[:, :k] = kmer_strings(kmers, k).view(np.uint8).reshape(U, k)
        body[:, k:roww - 1:2] = 9
        g = np.arange(G)
        bits = (matrix[g >> 6, :] >> (63 - (g & 63)).astype(np.uint64)[:, None]) & np.uint64(1)   # (G, U)
        body[:, k + 1:roww - 1:2] = (bits.T + ord("0")).astype(np.uint8)
        body[:, roww - 1] = 10
    return np.concatenate([np.frombuffer(header, dtype=np.uint8), body.reshape(-1)])


def write_tsv(path: str, kmers, matrix, names, k: int) -> None:
    format_tsv(kmers, matrix, names, k).tofile(path)
