"""Pooled k-mer count table on the GPU -- stand-in for the GUI's standalone DSK run
(``App.run_dsk``, src/app.py:1356-1416: ``ls -1 <folder>/*.fna > dsk_output`` then
``dsk -file dsk_output -out-dir <dir> -kmer-size <k>``).

DSK counts the canonical k-mers of ALL listed files together and keeps those whose abundance reaches
``-abundance-min`` (DSK's own default, 2, applies because the GUI passes none -- SURVEY.md A5 [UP]).  Here the same
table comes from one libgrmkm build with ``GRMKM_FLAG_COUNTS`` (every file on one row, per-k-mer counters in the
shared-memory tables, counts retained).  The real tool writes GATB's HDF5 layout, whose source is not in the
reference; this module writes a plain HDF5 file ``dsk_output.h5`` with ``kmers`` (uint64, canonical integers
A0 C1 T2 G3), ``counts`` (uint32) and ``kmer_sequences`` (S<k>) plus, on request, the ``dsk2ascii`` text form.

    python -m grm_b200.dsk -file LIST -out-dir DIR -kmer-size 31 [-abundance-min 2] [-ascii]
"""
from __future__ import annotations

import glob
import os
import sys

import numpy as np

from . import hdf5min
from .kover_cmd import to_linux_path
from .native import FASTA, FLAG_COUNTS

DSK_DEFAULT_ABUNDANCE_MIN = 2


def dsk_command(dsk_path: str, dataset_folder: str, output_directory: str, kmer_size=31) -> str:
    """The two shell lines App.run_dsk builds (src/app.py:1371-1372), with ``dsk_path`` free to point at this module."""
    config_path = f"{output_directory}/dsk_output"
    ls_command = f'ls -1 {to_linux_path(dataset_folder)}/*.fna > "{to_linux_path(config_path)}"'
    run = f'"{to_linux_path(dsk_path)}" -file "{to_linux_path(config_path)}" -out-dir "{to_linux_path(output_directory)}" -kmer-size {kmer_size}'
    return f"{ls_command}\n{run}"


def count_files(paths, kmer_size=31, abundance_min=DSK_DEFAULT_ABUNDANCE_MIN, input_kind=FASTA, device=-1):
    """-> (kmers uint64[U] in ascending hash order, counts uint32[U], stats)."""
    from .builder import KmerMatrixBuilder
    with KmerMatrixBuilder(k=int(kmer_size), min_abundance=max(1, int(abundance_min)), input_kind=input_kind,
                           flags=FLAG_COUNTS, device=device) as b:
        for p in paths:
            b.add_genome_files(0, [p])
        b.build()
        kmers, counts, seqs = b.kmers(), b.matrix(), b.kmer_strings()
        return kmers, counts.reshape(-1).astype(np.uint32), seqs, b.stats


def run_dsk(dataset_folder: str, output_directory: str, kmer_size=31, abundance_min=DSK_DEFAULT_ABUNDANCE_MIN,
            ascii_dump: bool = False, file_list=None) -> str:
    """Function form of App.run_dsk: pooled counts of every ``*.fna`` of the folder -> ``<out>/dsk_output.h5``."""
    if not kmer_size:
        kmer_size = 31
    paths = list(file_list) if file_list is not None else sorted(glob.glob(os.path.join(dataset_folder, "*.fna")))
    if not paths:
        raise FileNotFoundError("No .fna files found in the selected dataset folder.")
    os.makedirs(output_directory, exist_ok=True)
    kmers, counts, seqs, stats = count_files(paths, kmer_size, abundance_min)
    out = os.path.join(output_directory, "dsk_output.h5")
    with hdf5min.H5Writer(out) as h5:
        h5.attrs.update({"kmer_size": int(kmer_size), "abundance_min": int(abundance_min), "n_files": len(paths),
                         "n_bases": int(stats["n_bases"]), "n_kmer_occurrences": int(stats["n_windows"])})
        h5.create_dataset("kmers", kmers, gzip=4)
        h5.create_dataset("counts", counts, gzip=4)
        h5.create_dataset("kmer_sequences", seqs, gzip=4)
    if ascii_dump:
        with open(os.path.join(output_directory, "dsk_output.txt"), "wb") as f:
            for s, c in zip(seqs.tolist(), counts.tolist()):
                f.write(s + b" %d\n" % c)
    print("DSK (grm_b200): %d files, %d bases, %d distinct k-mers with abundance >= %d -> %s"
          % (len(paths), stats["n_bases"], len(kmers), int(abundance_min), out))
    return out


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    opts = {"-kmer-size": "31", "-abundance-min": str(DSK_DEFAULT_ABUNDANCE_MIN), "-out-dir": ".", "-file": None}
    ascii_dump = False
    i = 0
    while i < len(argv):
        a = argv[i]
        if a == "-ascii":
            ascii_dump = True
            i += 1
        elif a in opts and i + 1 < len(argv):
            opts[a] = argv[i + 1]
            i += 2
        else:
            print("usage: python -m grm_b200.dsk -file LIST -out-dir DIR [-kmer-size 31] [-abundance-min 2] [-ascii]\n"
                  "unsupported option: %s" % a, file=sys.stderr)
            return 2
    if not opts["-file"]:
        print("Error: -file is required", file=sys.stderr)
        return 2
    try:
        src = opts["-file"]
        if os.path.isfile(src) and not src.endswith((".fna", ".fa", ".fasta")):
            with open(src) as f:                       # a file of file names, one per line (what `ls -1` wrote)
                paths = [ln.strip() for ln in f if ln.strip()]
        else:
            paths = src.split(",")
        run_dsk(os.path.dirname(paths[0]) if paths else ".", opts["-out-dir"], int(opts["-kmer-size"]),
                int(opts["-abundance-min"]), ascii_dump, file_list=paths)
    except Exception as e:
        print("Error: %s" % e, file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
