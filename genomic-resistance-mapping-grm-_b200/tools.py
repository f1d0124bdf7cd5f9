"""Signature-compatible stand-ins for the reference's two native-tool shims
(``kover/dataset/tools/kmer_count.py:23-53`` and ``kmer_pack.py:23-39``).

The reference runs multidsk (one DSK count per genome, temp ``<basename>.h5`` files) and then dsk2kover
(merge + pack, appended to the HDF5 file).  The GPU path fuses both, so ``*_count_kmers`` only records
the job (a JSON manifest in ``out_dir``) and ``*_pack_kmers`` runs the fused build and adds
``kmer_sequences``, ``kmer_matrix`` and ``kmer_by_matrix_column`` to ``out_path``.
``grm_b200.create.from_contigs/from_reads`` call the builder directly and do not go through here.
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import hdf5min
from .create import BLOCK_SIZE, _minimum_uint_size
from .native import FASTA, FASTQ

_MANIFEST = "grmkm_count_manifest.json"


def _count(file_path, out_dir, kmer_size, abundance_min, out_compress, nb_cores, verbose, progress, kind):
    with open(file_path) as f:
        genomes = [line.strip().split(",") for line in f if line.strip()]
    manifest = {"k": int(kmer_size), "abundance_min": int(abundance_min), "kind": kind, "genomes": genomes}
    with open(os.path.join(str(out_dir), _MANIFEST), "w") as f:
        json.dump(manifest, f)


def contigs_count_kmers(file_path, out_dir, kmer_size, out_compress, nb_cores, verbose, progress):
    _count(file_path, out_dir, kmer_size, 1, out_compress, nb_cores, verbose, progress, FASTA)   # -abundance-min 1


def reads_count_kmers(file_path, out_dir, kmer_size, abundance_min, out_compress, nb_cores, verbose, progress):
    _count(file_path, out_dir, kmer_size, abundance_min, out_compress, nb_cores, verbose, progress, FASTQ)


def contigs_pack_kmers(file_path, out_path, filter_singleton, kmer_length, compression, chunk_size, nb_genomes, progress):
    """file_path = the list of per-genome count files (``list_h5``); its directory holds the manifest."""
    from .builder import KmerMatrixBuilder
    with open(os.path.join(os.path.dirname(str(file_path)), _MANIFEST)) as f:
        m = json.load(f)
    if int(kmer_length) != m["k"] or int(nb_genomes) != len(m["genomes"]):
        raise ValueError("pack step does not match the preceding count step")
    with KmerMatrixBuilder(k=m["k"], min_abundance=m["abundance_min"], keep_singletons=(filter_singleton == "nothing"),
                           input_kind=m["kind"]) as b:
        b.set_genome_count(len(m["genomes"]))
        for row, files in enumerate(m["genomes"]):
            b.add_genome_files(row, files)
        b.build()
        seqs, mat = b.kmer_strings(), b.matrix()
    old_attrs, old_ds = {}, []
    if os.path.exists(str(out_path)):
        r = hdf5min.H5Reader(str(out_path))
        old_attrs = r.attrs
        old_ds = [(n, d.read(), d.attrs, bool(d._filters)) for n, d in r.datasets.items()]
    gz = int(compression)
    U = mat.shape[1]
    with hdf5min.H5Writer(str(out_path)) as h5:
        h5.attrs.update(old_attrs)
        for n, data, attrs, was_gz in old_ds:
            if n not in ("kmer_sequences", "kmer_matrix", "kmer_by_matrix_column"):
                h5.create_dataset(n, data, gzip=gz if was_gz else 0, attrs=attrs)
        h5.create_dataset("kmer_sequences", seqs, gzip=gz)
        h5.create_dataset("kmer_matrix", mat, chunks=(1, max(1, min(U, int(chunk_size) or BLOCK_SIZE))), gzip=gz)
        h5.create_dataset("kmer_by_matrix_column", np.arange(U, dtype=_minimum_uint_size(U)), gzip=gz)


reads_pack_kmers = contigs_pack_kmers
