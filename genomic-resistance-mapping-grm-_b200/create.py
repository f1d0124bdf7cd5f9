"""Kover dataset creation on the GPU path -- Python-3 stand-in for the reference's
``bin/kover/core/kover/dataset/create.py`` (from_contigs :278-396, from_reads :399-523,
from_tsv :119-275, _parse_metadata :65-116) with the same signatures, callbacks and file layout.

Where the reference forks ``multidsk`` and ``dsk2kover`` (tools/kmer_count.py:28-37,
tools/kmer_pack.py:28-36) this module calls libgrmkm through :class:`KmerMatrixBuilder`;
the HDF5 file is written by :mod:`hdf5min` (no h5py in the image).  Differences, all supersets:
no phenotype is allowed for contigs/reads (the reference raises NameError, create.py:345);
native failures raise instead of being ignored (kmer_count.py:28); ``temp_dir`` is accepted but
unused (there are no per-genome temp files any more).
"""
from __future__ import annotations

import logging
import os
from math import ceil
from time import time
from uuid import uuid1

import numpy as np

from . import hdf5min
from .native import FASTA, FASTQ

KMER_MATRIX_PACKING_SIZE = 64
KMER_MATRIX_DTYPE = np.uint64
PHENOTYPE_LABEL_DTYPE = np.uint8
BLOCK_SIZE = 100000          # chunk width of kmer_matrix (create.py:41, 224-230)
SUPPORTED_READ_EXTENSIONS = (".fastq", ".fastq.gz")


def _minimum_uint_size(max_value):
    """Smallest unsigned dtype holding max_value (kover/utils.py:117-130)."""
    for t in (np.uint8, np.uint16, np.uint32, np.uint64):
        if max_value <= np.iinfo(t).max:
            return t
    raise ValueError("value does not fit an unsigned 64-bit integer")


def _pack_binary_bytes_to_ints(a, pack_size=64):
    """(G x n) 0/1 bytes -> (ceil(G/64) x n) uint64, genome g at bit 63-(g%64) of word g//64
    (same layout as kover/utils.py:133-156; vectorised with packbits instead of a row loop)."""
    if pack_size != 64:
        raise ValueError("Supported data types are 64-bit integers.")
    a = np.asarray(a, dtype=np.uint8)
    G, n = a.shape
    W = int(ceil(G / 64.0))
    padded = np.zeros((W * 64, n), dtype=np.uint8)
    padded[:G] = a != 0
    by = np.packbits(padded.reshape(W, 64, n), axis=1)           # (W, 8, n): byte 0 holds genomes 0..7, MSB first
    return np.ascontiguousarray(by.transpose(0, 2, 1)).view(">u8").reshape(W, n).astype(np.uint64)


def _callbacks(warning_callback, error_callback, progress_callback=None):
    if warning_callback is None:
        warning_callback = logging.warning
    if error_callback is None:
        def error_callback(exception):
            raise exception
    if progress_callback is None:
        def progress_callback(task, fraction):
            return None
    return warning_callback, error_callback, progress_callback


def _parse_metadata(metadata_path, matrix_genome_ids, warning_callback, error_callback):
    """``GENOME_ID<ws>LABEL`` lines -> (genome ids kept, uint8 labels, label tags, classification type).

    Labels are numbered by their sorted order; genomes are kept in METADATA order when they also
    have genomic data; the two one-sided differences only warn (pinned by tests/golden/metadata.json)."""
    logging.debug("Parsing metadata.")
    ids, labels = [], []
    with open(metadata_path, "r") as f:
        for line in f:
            parts = line.split()
            if not parts:
                continue
            ids.append(parts[0])
            labels.append(parts[1])
    tags = sorted(set(labels))
    if len(tags) < 2:
        error_callback(Exception("The dataset must contain at least 2 different phenotypes"))
    elif len(tags) > 255:
        error_callback(Exception("The dataset can contain at most 255 different phenotypes"))
    classification_type = "binary" if len(tags) == 2 else "multiclass"
    logging.debug("The dataset problem type is " + classification_type + " classification.")
    index = {t: i for i, t in enumerate(tags)}
    if len(ids) > len(set(ids)):
        error_callback(Exception("The metadata contains multiple values for the same genome."))
    have = set(matrix_genome_ids)
    only_matrix = have - set(ids)
    if only_matrix:
        warning_callback("Missing metadata for %d genomes (%s). These genomes will be discarded." % (
            len(only_matrix), ", ".join(sorted(only_matrix))))
    only_md = set(ids) - have
    if only_md:
        warning_callback("The metadata contains values for %d genomes that are not in the genomic data (%s)." % (
            len(only_md), ", ".join(sorted(only_md))))
    keep = [(g, index[l]) for g, l in zip(ids, labels) if g in have]
    return (np.array([g for g, _ in keep]), np.array([l for _, l in keep], dtype=np.uint8),
            np.array(tags), classification_type)


def _check_phenotype_args(phenotype_description, phenotype_metadata_path, error_callback):
    if (phenotype_description is None) != (phenotype_metadata_path is None):
        error_callback(ValueError("If a phenotype is specified, it must have a description and a metadata file."))


def _ordered_genomes(data_ids, phenotype_description, phenotype_metadata_path, warning_callback, error_callback):
    """-> (genome_ids in final row order, labels or None, tags, classification_type or None)."""
    if phenotype_description is None:
        return np.array(list(data_ids)), None, np.array([], dtype="S1"), None
    genome_ids, labels, tags, ctype = _parse_metadata(phenotype_metadata_path, list(data_ids), warning_callback,
                                                      error_callback)
    logging.debug("Sorting genomes by metadata label for optimal performance.")
    order = np.argsort(labels, kind="stable")          # reference: unstable argsort; row order within a label is free
    return genome_ids[order], labels[order], tags, ctype


def _write_dataset(output_path, root_attrs, genome_ids, labels, tags, phenotype_description, kmer_strings, matrix,
                   gzip, threads):
    U = int(matrix.shape[1]) if matrix.ndim == 2 else 0
    with hdf5min.H5Writer(output_path, threads=threads) as h5:
        h5.attrs.update(root_attrs)
        if labels is not None:
            h5.create_dataset("phenotype", labels.astype(PHENOTYPE_LABEL_DTYPE),
                              attrs={"description": phenotype_description})
        h5.create_dataset("genome_identifiers", np.asarray(genome_ids).astype("S"), gzip=gzip)
        h5.create_dataset("phenotype_tags", np.asarray(tags).astype("S") if len(tags) else np.zeros(0, dtype="S1"),
                          gzip=gzip)
        h5.create_dataset("kmer_sequences", kmer_strings, gzip=gzip)
        h5.create_dataset("kmer_matrix", matrix.astype(KMER_MATRIX_DTYPE, copy=False),
                          chunks=(1, max(1, min(U, BLOCK_SIZE))), gzip=gzip)
        h5.create_dataset("kmer_by_matrix_column", np.arange(U, dtype=_minimum_uint_size(U)), gzip=gzip)


def _root_attrs(source_type, genomic_data, phenotype_description, phenotype_metadata_path, gzip, ctype,
                filter_singleton=None):
    attrs = {
        "created": float(time()),
        "uuid": str(uuid1()),
        "genome_source_type": source_type,
        "genomic_data": str(genomic_data),
        "phenotype_description": phenotype_description if phenotype_description is not None else "NA",
        "phenotype_metadata_source": phenotype_metadata_path if phenotype_metadata_path is not None else "NA",
    }
    if filter_singleton is not None:
        attrs["filter"] = filter_singleton
    attrs["compression"] = "gzip (level %d)" % gzip
    if ctype is not None:
        attrs["classification_type"] = ctype
    return attrs


def _host_threads(nb_cores):
    try:
        n = int(nb_cores)
    except (TypeError, ValueError):
        n = 0
    return n if n > 0 else (os.cpu_count() or 1)


def _read_inputs(files_per_genome, threads):
    """Every input file in page-locked host memory (one arena, files 16-byte aligned), read by a thread pool: the
    H2D copies of the build then run at PCIe speed without a pageable staging hop.  ``.gz`` files are left to the
    library's own inflater.  -> (arena tensor or None, rows, buffers, leftover [(row, path)])"""
    from concurrent.futures import ThreadPoolExecutor
    flat = [(row, p) for row, files in enumerate(files_per_genome) for p in files]
    plain = [(row, p) for row, p in flat if not p.endswith(".gz")]
    rest = [(row, p) for row, p in flat if p.endswith(".gz")]
    sizes = [os.path.getsize(p) for _, p in plain]
    offs, total = [], 0
    for n in sizes:
        offs.append(total)
        total += (n + 15) & ~15
    if not total:
        return None, [], [], rest
    import torch
    try:
        arena = torch.empty(total, dtype=torch.uint8).pin_memory()
    except RuntimeError:                                   # not enough lockable memory: pageable works too, slower
        arena = torch.empty(total, dtype=torch.uint8)
    host = arena.numpy()

    def load(i):
        n, o = sizes[i], offs[i]
        with open(plain[i][1], "rb", buffering=0) as f:
            got = f.readinto(memoryview(host[o:o + n]))
            while got < n:                                  # readinto may stop short on some file systems
                more = f.readinto(memoryview(host[o + got:o + n]))
                if not more:
                    raise IOError("short read on %s" % plain[i][1])
                got += more

    with ThreadPoolExecutor(max_workers=max(1, min(threads, 32))) as ex:
        list(ex.map(load, range(len(plain))))
    return arena, [row for row, _ in plain], [host[o:o + n] for o, n in zip(offs, sizes)], rest


def _build(files_per_genome, kmer_size, abundance_min, filter_singleton, input_kind, progress, gpus=None,
           temp_dir=None, threads=8):
    from . import multi
    k = int(kmer_size)
    keep = (filter_singleton == "nothing")
    n_gpus = multi.requested_gpus(gpus)
    if n_gpus > 1:
        # GRM_GPUS / gpus=: one process per GPU, rows sharded in 64-aligned blocks (multi.py)
        if progress:
            print("grm_b200: counting and packing k-mers of %d genomes on %d GPUs" % (len(files_per_genome), n_gpus), flush=True)
        return multi.build_matrix(files_per_genome, k, max(1, int(abundance_min)), keep, input_kind, n_gpus, temp_dir)
    from .builder import KmerMatrixBuilder
    with KmerMatrixBuilder(k=k, min_abundance=max(1, int(abundance_min)), keep_singletons=keep,
                           input_kind=input_kind) as b:
        b.set_genome_count(len(files_per_genome))
        arena, rows, bufs, rest = _read_inputs(files_per_genome, threads)
        if rows:
            b.add_genomes(rows, bufs)
        for row, path in rest:
            b.add_genome_files(row, [path])
        if progress:
            print("grm_b200: counting and packing k-mers of %d genomes on the GPU" % len(files_per_genome), flush=True)
        b.build()
        del arena
        stats = b.stats
        logging.debug("k-mer matrix: %d columns, %d bases, device times %s", stats["n_kmers"], stats["n_bases"], b.times)
        _, mat = b.result_host()                     # one D2H at PCIe speed into the context's page-locked buffer
        return b.kmer_strings(), np.array(mat), stats


def from_contigs(contig_list_path, output_path, kmer_size, filter_singleton, phenotype_description,
                 phenotype_metadata_path, gzip, temp_dir, nb_cores, verbose, progress, warning_callback=None,
                 error_callback=None, gpus=None):
    """``kover dataset create from-contigs`` (kover:103-159 -> create.py:278-396)."""
    warning_callback, error_callback, _ = _callbacks(warning_callback, error_callback)
    gzip = int(gzip)
    _check_phenotype_args(phenotype_description, phenotype_metadata_path, error_callback)
    contig_file_by_genome_id = {}
    with open(contig_list_path, "r") as f:
        for line in f:
            parts = line.split()
            if parts:
                contig_file_by_genome_id[parts[0]] = parts[1]       # later duplicates win (dict(), create.py:302)
    for g_id, contig_file in contig_file_by_genome_id.items():
        if not os.path.exists(contig_file):
            error_callback(IOError("The contig file for genome %s cannot be found: %s" % (str(g_id), contig_file)))
    logging.debug("The k-mer matrix contains %d genomes." % len(contig_file_by_genome_id))
    genome_ids, labels, tags, ctype = _ordered_genomes(contig_file_by_genome_id.keys(), phenotype_description,
                                                       phenotype_metadata_path, warning_callback, error_callback)
    files = [[contig_file_by_genome_id[str(g)]] for g in genome_ids]
    logging.debug("Counting and packing k-mers (libgrmkm).")
    kmer_strings, matrix, _ = _build(files, kmer_size, 1, filter_singleton, FASTA, progress, gpus, temp_dir,
                                     _host_threads(nb_cores))                                   # -abundance-min 1
    attrs = _root_attrs("contigs", contig_list_path, phenotype_description, phenotype_metadata_path, gzip, ctype,
                        filter_singleton)
    _write_dataset(output_path, attrs, genome_ids, labels, tags, phenotype_description, kmer_strings, matrix, gzip,
                   _host_threads(nb_cores))
    logging.debug("Dataset creation completed.")


def from_reads(reads_folders_list_path, output_path, kmer_size, abundance_min, filter_singleton, phenotype_description,
               phenotype_metadata_path, gzip, temp_dir, nb_cores, verbose, progress, warning_callback=None,
               error_callback=None, gpus=None):
    """``kover dataset create from-reads`` (kover:161-224 -> create.py:399-523)."""
    warning_callback, error_callback, _ = _callbacks(warning_callback, error_callback)
    gzip = int(gzip)
    _check_phenotype_args(phenotype_description, phenotype_metadata_path, error_callback)
    reads_folder_by_genome_id = {}
    with open(reads_folders_list_path, "r") as f:
        for line in f:
            parts = line.split()
            if parts:
                reads_folder_by_genome_id[parts[0]] = parts[1]
    for g_id, read_dir in reads_folder_by_genome_id.items():
        if not os.path.exists(read_dir):
            error_callback(IOError("The read directory for genome %s cannot be found: %s" % (str(g_id), read_dir)))
    logging.debug("The k-mer matrix contains %d genomes." % len(reads_folder_by_genome_id))
    genome_ids, labels, tags, ctype = _ordered_genomes(reads_folder_by_genome_id.keys(), phenotype_description,
                                                       phenotype_metadata_path, warning_callback, error_callback)
    files = []
    for g in genome_ids:
        d = reads_folder_by_genome_id[str(g)]
        files.append([os.path.join(d, name) for name in sorted(os.listdir(d)) if name.endswith(SUPPORTED_READ_EXTENSIONS)])
    kmer_strings, matrix, _ = _build(files, kmer_size, abundance_min, filter_singleton, FASTQ, progress, gpus, temp_dir,
                                     _host_threads(nb_cores))
    attrs = _root_attrs("reads", reads_folders_list_path, phenotype_description, phenotype_metadata_path, gzip, ctype,
                        filter_singleton)
    _write_dataset(output_path, attrs, genome_ids, labels, tags, phenotype_description, kmer_strings, matrix, gzip,
                   _host_threads(nb_cores))
    logging.debug("Dataset creation completed.")


def read_kmer_matrix_tsv_rows(tsv_path):
    """Ray Surveyor KmerMatrix.tsv -> (genome ids, k, row width, body uint8[n][row width]) without touching the cells
    (the GPU packer reads them, grmkm_tsv_pack).  The body is a memory map of the file."""
    with open(tsv_path, "rb") as f:
        header = f.readline()
        hdr_len = f.tell()
        first = f.readline()
    genome_ids = header.rstrip(b"\r\n").decode().split("\t")[1:]
    size = os.path.getsize(tsv_path) - hdr_len
    if size == 0:
        return genome_ids, 0, 0, np.zeros((0, 0), dtype=np.uint8)
    roww = len(first)
    if size % roww != 0:
        raise Exception("The k-mer matrix rows do not all have the same width.")
    k = first.index(b"\t")
    if roww != k + 2 * len(genome_ids) + 1:
        raise Exception("Unexpected k-mer matrix row width.")
    body = np.memmap(tsv_path, dtype=np.uint8, mode="r", offset=hdr_len, shape=(size // roww, roww))
    return genome_ids, k, roww, body


def read_kmer_matrix_tsv(tsv_path):
    """Ray Surveyor KmerMatrix.tsv -> (genome ids, kmer strings S<k>[n], presence uint8[n][G]).

    Grammar enforced by the reference's consumer (create.py:121-137): header ``kmers\\t<id>...``,
    then fixed-width rows ``<kmer>\\t<0|1>...``; (size - header) must be a multiple of the row width."""
    with open(tsv_path, "rb") as f:
        header = f.readline()
        body = np.frombuffer(f.read(), dtype=np.uint8)
    genome_ids = header.rstrip(b"\r\n").decode().split("\t")[1:]
    if body.size == 0:
        return genome_ids, np.zeros(0, dtype="S1"), np.zeros((0, len(genome_ids)), dtype=np.uint8)
    first_nl = int(np.argmax(body == 10))
    roww = first_nl + 1
    if body.size % roww != 0:
        raise Exception("The k-mer matrix rows do not all have the same width.")
    rows = body.reshape(-1, roww)
    k = int(np.argmax(rows[0] == 9))
    if roww != k + 2 * len(genome_ids) + 1:
        raise Exception("Unexpected k-mer matrix row width.")
    kmers = np.ascontiguousarray(rows[:, :k]).view(f"S{k}").reshape(-1)
    cells = rows[:, k + 1:roww - 1:2] - ord("0")
    return genome_ids, kmers, cells


def from_tsv(tsv_path, output_path, phenotype_description, phenotype_metadata_path, gzip, warning_callback=None,
             error_callback=None, progress_callback=None, use_gpu=None):
    """``kover dataset create from-tsv`` (kover:41-100 -> create.py:119-275).  The reference parses the text with pandas
    in chunks and packs the rows in a Python loop (create.py:241-271, utils.py:147-154); here the fixed-width rows go to
    the GPU as they are and one kernel gathers the cells of the kept genomes into matrix words (grmkm_tsv_pack).
    ``use_gpu=False`` (or GRM_TSV_PACKER=host) selects the numpy packer instead -- an explicit choice, never a
    fallback: with the GPU packer selected and no device the call fails."""
    if use_gpu is None:
        use_gpu = os.environ.get("GRM_TSV_PACKER", "gpu") != "host"
    if use_gpu:
        return _from_tsv_gpu(tsv_path, output_path, phenotype_description, phenotype_metadata_path, gzip, warning_callback,
                             error_callback, progress_callback)
    warning_callback, error_callback, progress_callback = _callbacks(warning_callback, error_callback, progress_callback)
    gzip = int(gzip)
    if (phenotype_description is None) != (phenotype_metadata_path is None):
        raise ValueError("If a phenotype is specified, it must have a description and a metadata file.")
    progress_callback("Creating", 0.0)
    tsv_ids, kmers, cells = read_kmer_matrix_tsv(tsv_path)
    logging.debug("The k-mer matrix contains %d genomes." % len(tsv_ids))
    if len(set(tsv_ids)) < len(tsv_ids):
        error_callback(Exception("The genomic data contains genomes with the same identifier."))
    genome_ids, labels, tags, ctype = _ordered_genomes(tsv_ids, phenotype_description, phenotype_metadata_path,
                                                       warning_callback, error_callback)
    col = {g: i for i, g in enumerate(tsv_ids)}
    sel = np.array([col[str(g)] for g in genome_ids], dtype=np.int64)
    progress_callback("Creating", 0.5)
    matrix = _pack_binary_bytes_to_ints(np.ascontiguousarray(cells[:, sel].T) if cells.size else
                                        np.zeros((len(sel), 0), dtype=np.uint8), KMER_MATRIX_PACKING_SIZE)
    attrs = _root_attrs("tsv", tsv_path, phenotype_description, phenotype_metadata_path, gzip, ctype)
    _write_dataset(output_path, attrs, genome_ids, labels, tags, phenotype_description, kmers, matrix, gzip,
                   os.cpu_count() or 1)
    progress_callback("Creating", 1.0)
    logging.debug("Dataset creation completed.")


def _from_tsv_gpu(tsv_path, output_path, phenotype_description, phenotype_metadata_path, gzip, warning_callback,
                  error_callback, progress_callback):
    from .builder import KmerMatrixBuilder
    warning_callback, error_callback, progress_callback = _callbacks(warning_callback, error_callback, progress_callback)
    gzip = int(gzip)
    if (phenotype_description is None) != (phenotype_metadata_path is None):
        raise ValueError("If a phenotype is specified, it must have a description and a metadata file.")
    progress_callback("Creating", 0.0)
    tsv_ids, k, roww, body = read_kmer_matrix_tsv_rows(tsv_path)
    logging.debug("The k-mer matrix contains %d genomes." % len(tsv_ids))
    if len(set(tsv_ids)) < len(tsv_ids):
        error_callback(Exception("The genomic data contains genomes with the same identifier."))
    genome_ids, labels, tags, ctype = _ordered_genomes(tsv_ids, phenotype_description, phenotype_metadata_path,
                                                       warning_callback, error_callback)
    col = {g: i for i, g in enumerate(tsv_ids)}
    sel = np.array([col[str(g)] for g in genome_ids], dtype=np.uint32)
    progress_callback("Creating", 0.5)
    n = int(body.shape[0])
    if n:
        kmers = np.ascontiguousarray(body[:, :k]).view("S%d" % k).reshape(-1)
        with KmerMatrixBuilder(k=min(max(k, 1), 32)) as b:           # any context of the device will do
            matrix = b.tsv_pack(body, roww, k, len(tsv_ids), sel)
    else:
        kmers = np.zeros(0, dtype="S1")
        matrix = np.zeros(((len(sel) + 63) // 64, 0), dtype=np.uint64)
    attrs = _root_attrs("tsv", tsv_path, phenotype_description, phenotype_metadata_path, gzip, ctype)
    _write_dataset(output_path, attrs, genome_ids, labels, tags, phenotype_description, kmers, matrix, gzip,
                   os.cpu_count() or 1)
    progress_callback("Creating", 1.0)
    logging.debug("Dataset creation completed.")
