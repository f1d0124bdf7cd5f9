"""``kover``-compatible command line for the dataset-creation path, so that the GUI's
``create_command(kover_path=...)`` (src/kover.py:52-108) can point at it unchanged:

    python -m grm_b200.cli dataset create from-contigs --genomic-data L --output O [--kmer-size 31] ...

Flags, defaults and the ``filter`` mapping follow bin/kover/kover:36-224.  Unlike the reference the
exit status is non-zero on failure (the GUI treats non-zero as failure, src/app.py:3406).
Only ``dataset create`` exists here; split / info / learn are out of scope.
"""
from __future__ import annotations

import argparse
import logging
import sys
from tempfile import gettempdir

VERSION = "grm_b200 0.1.0 (kover dataset create compatible)"


def _common(parser, native: bool):
    parser.add_argument("--genomic-data", required=True)
    parser.add_argument("--phenotype-description")
    parser.add_argument("--phenotype-metadata")
    parser.add_argument("--output", required=True)
    if native:
        parser.add_argument("--kmer-size", default=31, help="The k-mer size (max is 32 here; the reference allows 128).")
        parser.add_argument("--singleton-kmers", default=False, action="store_true")
        parser.add_argument("--n-cpu", "--n-cores", default=0)
        parser.add_argument("--temp-dir", default=gettempdir())
    parser.add_argument("--compression", type=int, default=4)
    parser.add_argument("-x", "--progress", action="store_true")
    parser.add_argument("-v", "--verbose", default=False, action="store_true")


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if argv[:1] == ["--version"]:
        print(VERSION)
        return 0
    if len(argv) < 3 or argv[0] != "dataset" or argv[1] != "create" or argv[2] not in ("from-tsv", "from-contigs", "from-reads"):
        print("usage: grm-kover dataset create {from-tsv,from-contigs,from-reads} [options]", file=sys.stderr)
        return 2
    source = argv[2]
    parser = argparse.ArgumentParser(prog="kover dataset create " + source,
                                     description="Creates a Kover dataset from genomic data and optionally phenotypic metadata")
    _common(parser, native=(source != "from-tsv"))
    if source == "from-reads":
        parser.add_argument("--kmer-min-abundance", default=1)
    if len(argv) == 3:
        argv.append("--help")
    args = parser.parse_args(argv[3:])
    if (args.phenotype_description is None) != (args.phenotype_metadata is None):
        print("Error: The phenotype description and metadata file must be specified.")
        return 1
    if args.verbose:
        logging.basicConfig(level=logging.DEBUG,
                            format="%(asctime)s.%(msecs)d %(levelname)s %(module)s - %(funcName)s: %(message)s")
    from . import create
    try:
        if source == "from-tsv":
            create.from_tsv(tsv_path=args.genomic_data, output_path=args.output,
                            phenotype_description=args.phenotype_description,
                            phenotype_metadata_path=args.phenotype_metadata, gzip=args.compression)
        else:
            filter_option = "nothing" if args.singleton_kmers else "singleton"
            kw = dict(output_path=args.output, kmer_size=args.kmer_size, filter_singleton=filter_option,
                      phenotype_description=args.phenotype_description, phenotype_metadata_path=args.phenotype_metadata,
                      gzip=args.compression, temp_dir=args.temp_dir, nb_cores=args.n_cpu, verbose=args.verbose,
                      progress=args.progress)
            if source == "from-contigs":
                create.from_contigs(contig_list_path=args.genomic_data, **kw)
            else:
                create.from_reads(reads_folders_list_path=args.genomic_data, abundance_min=args.kmer_min_abundance, **kw)
    except Exception as e:
        print("Error: %s" % e, file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
