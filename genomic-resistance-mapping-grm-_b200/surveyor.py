"""Ray Surveyor stand-in: ``survey.conf`` in, ``Surveyor/KmerMatrix.tsv`` out.

Replaces ``mpiexec -n 4 Ray survey.conf`` (src/app.py:1310).  ``generate_survey_conf`` is the function
form of ``App.generate_survey_conf`` (src/app.py:3812-3835, pinned by tests/golden/commands.json);
``parse_survey_conf`` accepts exactly the five directives that generator writes.  The matrix is the
presence matrix with min abundance 1 and no singleton filter; the TSV is the
fixed-width grammar ``kover dataset create from-tsv`` parses (create.py:121-137, 241-264).

    python -m grm_b200.surveyor survey.conf                          one GPU
    GRM_GPUS=N python -m grm_b200.surveyor survey.conf               rows sharded over N GPUs (spawns N ranks)
    torchrun --nproc-per-node N -m grm_b200.surveyor survey.conf     the same, joining an existing launch
"""
from __future__ import annotations

import os
import pathlib
import shlex
import sys

from .kover_cmd import to_linux_path


def generate_survey_conf(input_files, kmer_size, output_dir) -> str:
    path = os.path.join(output_dir, "survey.conf")
    lines = ["-k %s" % kmer_size, "-run-surveyor", "-output %s/survey.res" % to_linux_path(output_dir),
             "-write-kmer-matrix"]
    for f in input_files:
        lines.append("-read-sample-assembly %s %s" % (pathlib.Path(f).stem, to_linux_path(f)))
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return path


def parse_survey_conf(path) -> dict:
    conf = {"k": 31, "run_surveyor": False, "output": None, "write_kmer_matrix": False, "samples": []}
    with open(path) as f:
        for raw in f:
            line = raw.strip()
            if not line:
                continue
            tok = shlex.split(line)
            if tok[0] == "-k":
                conf["k"] = int(tok[1])
            elif tok[0] == "-run-surveyor":
                conf["run_surveyor"] = True
            elif tok[0] == "-output":
                conf["output"] = tok[1]
            elif tok[0] == "-write-kmer-matrix":
                conf["write_kmer_matrix"] = True
            elif tok[0] == "-read-sample-assembly":
                if len(tok) < 3:
                    raise ValueError("bad -read-sample-assembly line: %r" % line)
                # the generator does not quote sample names, so a name with spaces spans several tokens
                conf["samples"].append((" ".join(tok[1:-1]), tok[-1]))
            else:
                raise ValueError("unsupported Ray directive %r (only the Surveyor k-mer matrix path is implemented)" % tok[0])
    if conf["output"] is None:
        raise ValueError("survey.conf has no -output directive")
    return conf


def run_surveyor(conf_path, device: int = -1, gpus=None) -> str | None:
    """Build the matrix and write <output>/Surveyor/KmerMatrix.tsv.  Returns the path.  GRM_GPUS=N (or gpus=N): the
    rows are sharded over N GPUs, one process each (the GUI runs Ray as ``mpiexec -n 4``, src/app.py:1310), and every
    rank writes the rows of its own column slice into the file (multi.py, DistributedBuilder.write_tsv)."""
    from . import multi
    conf = parse_survey_conf(conf_path)
    names = [n for n, _ in conf["samples"]]
    out_dir = os.path.join(conf["output"], "Surveyor")
    out_path = os.path.join(out_dir, "KmerMatrix.tsv")
    n_gpus = multi.requested_gpus(gpus)
    under_torchrun = int(os.environ.get("WORLD_SIZE", "1")) > 1
    if n_gpus > 1 or under_torchrun:
        os.makedirs(out_dir, exist_ok=True)
        job = {"files": [[p] for _, p in conf["samples"]], "k": conf["k"], "min_abundance": 1, "keep_singletons": True,
               "input_kind": 0, "tsv": out_path, "names": names}
        multi.cleanup(multi.launch(job, n_gpus))
        return out_path
    from .builder import KmerMatrixBuilder
    with KmerMatrixBuilder(k=conf["k"], min_abundance=1, keep_singletons=True, device=device) as b:
        b.set_genome_count(len(names))
        for row, (_, path) in enumerate(conf["samples"]):
            b.add_genome_files(row, [path])
        b.build()
        os.makedirs(out_dir, exist_ok=True)
        b.tsv(names).tofile(out_path)
        print("[Surveyor] %d samples, %d k-mers -> %s" % (len(names), b.dims[0], out_path))
    return out_path


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        print("usage: python -m grm_b200.surveyor survey.conf", file=sys.stderr)
        return 2
    try:
        run_surveyor(argv[0])
    except Exception as e:  # the GUI only looks at the exit status (src/app.py:1322-1354)
        print("Error: %s" % e, file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
