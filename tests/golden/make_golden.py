#!/usr/bin/env python
"""Generate tests/golden/*.json by EXECUTING the reference's own Python code.

Run in the build container only (needs /root/reference; the GPU box does not
have it -- tests read the committed JSON, never the reference):

    python tests/golden/make_golden.py

The reference's Kover package is Python 2 and imports h5py, so it cannot be
imported as a module.  Instead each function is lifted out of its source file
with ``ast`` and exec'd unchanged in a Python 3 namespace with two shims
(``xrange = range``; integer ``ceil``).  Nothing is copied into the repo: the
JSON holds only inputs and the outputs the reference code produced.

Functions executed (path under /root/reference : function):
  bin/kover/core/kover/utils.py : _pack_binary_bytes_to_ints, _unpack_binary_bytes_from_ints,
                                  _minimum_uint_size
  bin/kover/core/kover/dataset/create.py : _parse_metadata
  src/util.py  : to_linux_path, quote_space
  src/kover.py : create_command, create_contigs_path_tsv (whole module, util stubbed with the two above)
  src/app.py   : App.generate_survey_conf
"""
import ast
import json
import logging
import os
import pathlib as pl
import re
import sys
import tempfile
import types
from math import ceil

import numpy as np

REF = os.environ.get("GRM_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def lift(path, names, namespace, class_name=None):
    """exec the named top-level (or class-level) function defs of ``path`` into namespace."""
    src = open(os.path.join(REF, path), encoding="utf-8").read()
    tree = ast.parse(src)
    body = tree.body
    if class_name:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name).body
    found = []
    for node in body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.decorator_list = []
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, "exec"), namespace)
            found.append(node.name)
    missing = set(names) - set(found)
    if missing:
        raise RuntimeError(f"{path}: functions not found: {missing}")
    return namespace


def golden_bits():
    ns = {"np": np, "ceil": ceil, "xrange": range}
    lift("bin/kover/core/kover/utils.py",
         ["_pack_binary_bytes_to_ints", "_unpack_binary_bytes_from_ints", "_minimum_uint_size"], ns)
    pack, unpack, minu = ns["_pack_binary_bytes_to_ints"], ns["_unpack_binary_bytes_from_ints"], ns["_minimum_uint_size"]
    cases = []
    fixed = [
        ("K1", np.array([[1, 0], [0, 1], [1, 1]], dtype=np.uint8)),
        ("K2", np.array([[1 if g in (0, 63, 64) else 0] for g in range(65)], dtype=np.uint8)),
    ]
    rng = np.random.RandomState(0x47524D31 & 0x7FFFFFFF)
    for G, n in [(1, 5), (2, 9), (63, 7), (64, 7), (65, 7), (100, 11), (128, 3), (130, 13), (200, 4)]:
        fixed.append((f"rand_G{G}_n{n}", (rng.rand(G, n) < 0.4).astype(np.uint8)))
    for name, a in fixed:
        b = pack(a, 64)
        u = unpack(b)
        assert (u[:a.shape[0]] == a).all() and not u[a.shape[0]:].any()
        cases.append({"name": name, "rows": a.tolist(), "packed_hex": [[format(int(x), "016x") for x in row] for row in b]})
    sizes = []
    for v in [0, 1, 255, 256, 65535, 65536, 2 ** 32 - 1, 2 ** 32, 2 ** 40, 2 ** 64 - 1]:
        sizes.append({"value": str(v), "dtype": np.dtype(minu(v)).name})
    return {"pack": cases, "minimum_uint_size": sizes}


def golden_metadata():
    ns = {"np": np, "xrange": range, "logging": logging}
    lift("bin/kover/core/kover/dataset/create.py", ["_parse_metadata"], ns)
    pm = ns["_parse_metadata"]
    cases = []
    specs = [
        ("binary_RS", "g1\tR\ng2\tS\ng3\tR\ng4\tS\n", ["g1", "g2", "g3", "g4"]),
        ("binary_01", "g1\t1\ng2\t0\ng3\t1\n", ["g3", "g1", "g2"]),
        ("binary_10_words", "a 1\nb 0\nc 0\nd 1\n", ["a", "b", "c", "d"]),
        ("multiclass", "s1\tresistant\ns2\tintermediate\ns3\tsusceptible\ns4\tresistant\n", ["s1", "s2", "s3", "s4"]),
        ("numeric_multiclass", "s1\t2\ns2\t0\ns3\t1\ns4\t10\n", ["s1", "s2", "s3", "s4"]),
        ("missing_metadata", "g1\tR\ng2\tS\n", ["g1", "g2", "g3"]),
        ("extra_metadata", "g1\tR\ng2\tS\ng9\tS\n", ["g2", "g1"]),
        ("order_from_metadata", "z\tS\ny\tR\nx\tS\nw\tR\n", ["w", "x", "y", "z"]),
    ]
    for name, text, ids in specs:
        warnings = []
        with tempfile.NamedTemporaryFile("w", suffix=".tsv", delete=False) as f:
            f.write(text)
        try:
            keep_ids, labels, tags, ctype = pm(f.name, ids, warnings.append, lambda e: (_ for _ in ()).throw(e))
        finally:
            os.unlink(f.name)
        cases.append({"name": name, "metadata": text, "matrix_genome_ids": ids,
                      "genome_ids": [str(x) for x in keep_ids], "labels": [int(x) for x in labels],
                      "tags": [str(x) for x in tags], "classification_type": ctype,
                      "n_warnings": len(warnings)})
    errs = []
    for name, text, ids in [("one_label", "g1\tR\ng2\tR\n", ["g1", "g2"]),
                            ("dup_genome", "g1\tR\ng1\tS\ng2\tS\n", ["g1", "g2"])]:
        with tempfile.NamedTemporaryFile("w", suffix=".tsv", delete=False) as f:
            f.write(text)
        got = []
        try:
            try:
                pm(f.name, ids, lambda w: None, lambda e: got.append(str(e)))
            except Exception as e:  # the reference falls through after error_callback returns
                got.append("raised:" + type(e).__name__)
        finally:
            os.unlink(f.name)
        errs.append({"name": name, "metadata": text, "matrix_genome_ids": ids, "first_error": got[0] if got else None})
    return {"cases": cases, "errors": errs}


def golden_commands():
    uns = {"os": os, "re": re}
    lift("src/util.py", ["to_linux_path", "quote_space"], uns)
    util = types.ModuleType("util")
    util.to_linux_path = uns["to_linux_path"]
    util.quote_space = uns["quote_space"]
    sys.modules["util"] = util
    kns = {"__name__": "ref_kover"}
    exec(compile(open(os.path.join(REF, "src/kover.py"), encoding="utf-8").read(), "src/kover.py", "exec"), kns)
    cc, Source = kns["create_command"], kns["Source"]
    out = {"to_linux_path": [], "create_command": [], "survey_conf": [], "contigs_path_tsv": None}
    for p in ["/data/genomes", "/data/my genomes/x.fna", "/Data/Upper", "/mnt/c/Users/a b/c"]:
        out["to_linux_path"].append({"in": p, "out": util.to_linux_path(p)})
    combos = [
        dict(source=Source.CONTIGS, genomic_data="/d/g_paths.tsv", output="/d/out/DATASET.kover",
             phenotype_description="/d/desc.txt", phenotype_metadata="/d/meta.tsv", kmer_size=31,
             kmer_min_abundance=1, singleton_kmers=False, n_cpu=4, compression=4, temp_dir="/tmp/t", x=True, v=False),
        dict(source=Source.CONTIGS, genomic_data="/d/g paths.tsv", output="/d/o.kover",
             phenotype_description=None, phenotype_metadata=None, kmer_size="21",
             kmer_min_abundance=5, singleton_kmers=True, n_cpu=0, compression=0, temp_dir="", x=False, v=True),
        dict(source=Source.READS, genomic_data="/d/reads.tsv", output="/d/o.kover",
             phenotype_description="/d/desc.txt", phenotype_metadata="/d/meta.tsv", kmer_size=15,
             kmer_min_abundance=2, singleton_kmers=True, n_cpu="8", compression=9, temp_dir=None, x=True, v=True),
        dict(source=Source.READS, genomic_data="/d/reads.tsv", output="/d/o.kover", kmer_min_abundance=0),
        dict(source=Source.K_MER_MATREX, genomic_data="/d/KmerMatrix.tsv", output="/d/o.kover",
             phenotype_description="/d/desc.txt", phenotype_metadata="/d/meta.tsv", kmer_size=31,
             kmer_min_abundance=3, singleton_kmers=True, n_cpu=4, compression=4, temp_dir="/tmp/t", x=False, v=False),
        dict(source=Source.CONTIGS, genomic_data="/d/g.tsv", output="/d/o.kover"),
    ]
    for kw in combos:
        out["create_command"].append({"kwargs": {k: (str(v) if k == "source" else v) for k, v in kw.items()},
                                      "command": cc("/opt/kover/bin/kover", **kw)})
    # create_contigs_path_tsv on a temp tree
    with tempfile.TemporaryDirectory() as td:
        gdir = os.path.join(td, "ecoli")
        os.makedirs(gdir)
        for n in ["562.100.fna", "562.200.fna", "notes.txt", "562.300.fa"]:
            open(os.path.join(gdir, n), "w").close()
        kns["create_contigs_path_tsv"](td, "ecoli")
        lines = open(gdir + "_paths.tsv", encoding="utf-8").read().replace(td, "<ROOT>")
        out["contigs_path_tsv"] = {"files": ["562.100.fna", "562.200.fna", "notes.txt", "562.300.fa"],
                                   "genome_name": "ecoli", "lines_sorted": sorted(lines.splitlines())}
    # App.generate_survey_conf
    ans = {"os": os, "pl": pl, "util": util, "tk": types.SimpleNamespace(W="w"), "ctk": None}
    lift("src/app.py", ["generate_survey_conf"], ans, class_name="App")
    gen = ans["generate_survey_conf"]
    with tempfile.TemporaryDirectory() as td:
        files = ["/data/ds/562.1.fna", "/data/ds/my genome.fna", "/data/ds/b.fna"]
        for k in (31, "21"):
            path = gen(None, files, k, td)
            text = open(path).read().replace(util.to_linux_path(td), "<OUT>")
            out["survey_conf"].append({"input_files": files, "kmer_size": k, "basename": os.path.basename(path), "text": text})
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit(f"reference not found at {REF} (only available in the build container)")
    for name, fn in [("bits.json", golden_bits), ("metadata.json", golden_metadata), ("commands.json", golden_commands)]:
        data = fn()
        data["_generated_by"] = "tests/golden/make_golden.py executing reference code (see docstring)"
        with open(os.path.join(OUT, name), "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
        print("wrote", name)


if __name__ == "__main__":
    main()
