"""CPU: the host-side mirror of the reference interface (command builders, metadata, survey.conf,
bit packer, from_tsv, CLI argument surface) against the golden vectors produced by executing the
reference's own Python code (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from grm_b200 import cli, create, hdf5min, kover_cmd, surveyor, tsv

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


# ---- src/kover.py, src/util.py, src/app.py:3812-3835 ---------------------------------------------
def test_create_command_matches_reference():
    for case in load("commands.json")["create_command"]:
        kw = dict(case["kwargs"])
        kw.setdefault("kover_path", "/opt/kover/bin/kover")
        assert kover_cmd.create_command(**kw) == case["command"], case


def test_to_linux_path_matches_reference():
    for c in load("commands.json")["to_linux_path"]:
        assert kover_cmd.to_linux_path(c["in"]) == c["out"]


def test_create_contigs_path_tsv_matches_reference(tmp_path):
    c = load("commands.json")["contigs_path_tsv"]
    d = tmp_path / c["genome_name"]
    d.mkdir()
    for name in c["files"]:
        (d / name).write_text(">x\nACGT\n")
    out = kover_cmd.create_contigs_path_tsv(str(tmp_path), c["genome_name"])
    assert out == str(tmp_path / (c["genome_name"] + "_paths.tsv"))
    got = sorted(open(out).read().splitlines())
    want = sorted(l.replace("<ROOT>", str(tmp_path)) for l in c["lines_sorted"])
    assert got == want


def test_generate_survey_conf_matches_reference_and_round_trips(tmp_path):
    for c in load("commands.json")["survey_conf"]:
        path = surveyor.generate_survey_conf(c["input_files"], c["kmer_size"], str(tmp_path))
        assert os.path.basename(path) == c["basename"]
        assert open(path).read() == c["text"].replace("<OUT>", str(tmp_path))
        conf = surveyor.parse_survey_conf(path)
        assert conf["k"] == int(c["kmer_size"]) and conf["run_surveyor"] and conf["write_kmer_matrix"]
        assert conf["output"] == str(tmp_path) + "/survey.res"
        assert [p for _, p in conf["samples"]] == c["input_files"]
        assert [n for n, _ in conf["samples"]] == [os.path.splitext(os.path.basename(f))[0] for f in c["input_files"]]


def test_survey_conf_rejects_other_directives(tmp_path):
    p = tmp_path / "survey.conf"
    p.write_text("-k 31\n-output /x\n-p reads_1.fastq reads_2.fastq\n")
    with pytest.raises(ValueError):
        surveyor.parse_survey_conf(str(p))
    p.write_text("-k 31\n-run-surveyor\n")
    with pytest.raises(ValueError):
        surveyor.parse_survey_conf(str(p))


# ---- kover/dataset/create.py:65-116, kover/utils.py:117-187 ---------------------------------------
def test_parse_metadata_matches_reference(tmp_path):
    g = load("metadata.json")
    for case in g["cases"]:
        p = tmp_path / (case["name"] + ".tsv")
        p.write_text(case["metadata"])
        warnings = []
        ids, labels, tags, ctype = create._parse_metadata(str(p), case["matrix_genome_ids"], warnings.append, _raise)
        assert list(ids) == case["genome_ids"], case["name"]
        assert labels.tolist() == case["labels"] and labels.dtype == np.uint8
        assert list(tags) == case["tags"] and ctype == case["classification_type"]
        assert len(warnings) == case["n_warnings"], (case["name"], warnings)
    for case in g["errors"]:
        p = tmp_path / (case["name"] + ".tsv")
        p.write_text(case["metadata"])
        errors = []
        try:
            create._parse_metadata(str(p), case["matrix_genome_ids"], lambda w: None, lambda e: errors.append(str(e)))
        except Exception:
            pass          # the reference keeps going after error_callback and may crash later; only the first report is pinned
        assert errors and errors[0] == case["first_error"], case["name"]


def _raise(e):
    raise e


def test_vectorised_packer_matches_reference_goldens():
    for case in load("bits.json")["pack"]:
        rows = np.array(case["rows"], dtype=np.uint8)
        want = np.array([[int(x, 16) for x in r] for r in case["packed_hex"]], dtype=np.uint64)
        got = create._pack_binary_bytes_to_ints(rows, 64)
        assert got.dtype == np.uint64 and np.array_equal(got, want), case["name"]
    with pytest.raises(ValueError):
        create._pack_binary_bytes_to_ints(np.zeros((2, 2), dtype=np.uint8), 32)


def test_minimum_uint_size_matches_reference():
    for c in load("bits.json")["minimum_uint_size"]:
        assert np.dtype(create._minimum_uint_size(int(c["value"]))).name == c["dtype"]


@pytest.fixture(autouse=True)
def _host_tsv_packer(monkeypatch):
    """These tests run without a GPU: from_tsv's numpy packer is selected explicitly (the GPU packer has its own -m gpu tests)."""
    monkeypatch.setenv("GRM_TSV_PACKER", "host")


# ---- Ray TSV text <-> matrix -------------------------------------------------------------------------
def _random_matrix(rng, G, U, k):
    kmers = np.sort(rng.choice(1 << min(62, 2 * k), size=U, replace=False).astype(np.uint64))
    present = (rng.random((G, U)) < 0.4).astype(np.uint8)
    present[rng.integers(G, size=U), np.arange(U)] = 1
    return kmers, present, create._pack_binary_bytes_to_ints(present, 64)


def test_host_tsv_writer_agrees_with_oracle_text():
    from oracle import oracle
    rng = np.random.default_rng(5)
    for G, U, k in [(1, 7, 5), (3, 50, 12), (64, 33, 31), (70, 129, 21)]:
        kmers, _, mat = _random_matrix(rng, G, U, k)
        names = [f"s{i}" for i in range(G)]
        got = tsv.format_tsv(kmers, mat, names, k).tobytes()
        assert got == oracle.py_format_tsv(kmers, mat, names, k)
        assert [s.decode() for s in tsv.kmer_strings(kmers, k)] == [oracle.py_kmer_string(int(x), k) for x in kmers]


@pytest.mark.parametrize("gzip", [0, 4])
def test_from_tsv_writes_the_kover_layout(tmp_path, gzip):
    """create.py:119-275 + consumer contract ds.py:26-148 (names, attrs, bit layout, chunks, dtypes)."""
    rng = np.random.default_rng(11)
    G, U, k = 67, 300, 15
    kmers, present, mat = _random_matrix(rng, G, U, k)
    names = [f"genome_{i}" for i in range(G)]
    tsv_path = tmp_path / "KmerMatrix.tsv"
    tsv.write_tsv(str(tsv_path), kmers, mat, names, k)
    # metadata: drop two genomes, one extra without data, labels R/S
    meta = tmp_path / "meta.tsv"
    keep = [n for i, n in enumerate(names) if i not in (3, 40)]
    labels = {n: ("R" if rng.random() < 0.5 else "S") for n in keep}
    meta.write_text("".join(f"{n}\t{labels[n]}\n" for n in keep) + "ghost\tR\n")
    desc = "resistance to something"
    out = tmp_path / "DATASET.kover"
    warnings, progress = [], []
    create.from_tsv(str(tsv_path), str(out), desc, str(meta), gzip, warnings.append, None,
                    lambda task, frac: progress.append((task, frac)))
    assert len(warnings) == 2 and progress[0] == ("Creating", 0.0) and progress[-1] == ("Creating", 1.0)
    r = hdf5min.H5Reader(str(out))
    a = r.attrs
    assert a["genome_source_type"] == "tsv" and a["genomic_data"] == str(tsv_path)
    assert a["phenotype_description"] == desc and a["phenotype_metadata_source"] == str(meta)
    assert a["compression"] == "gzip (level %d)" % gzip and a["classification_type"] == "binary"
    assert "filter" not in a and isinstance(a["created"], float) and len(a["uuid"]) == 36
    ids = [x.decode() for x in r["genome_identifiers"].read()]
    pheno = r["phenotype"].read()
    assert sorted(ids) == sorted(keep) and pheno.dtype == np.uint8
    assert pheno.tolist() == sorted(pheno.tolist())                      # genomes ordered by label
    assert [x.decode() for x in r["phenotype_tags"].read()] == ["R", "S"]
    assert all(labels[i] == "RS"[p] for i, p in zip(ids, pheno))
    assert r["phenotype"].attrs["description"] == desc
    seqs = r["kmer_sequences"].read()
    assert seqs.dtype == np.dtype(f"S{k}") and np.array_equal(seqs, tsv.kmer_strings(kmers, k))
    km = r["kmer_matrix"]
    assert km.dtype == np.uint64 and km.shape == ((len(keep) + 63) // 64, U) and km.chunks == (1, U)
    row_of = {n: i for i, n in enumerate(names)}
    want = create._pack_binary_bytes_to_ints(present[[row_of[i] for i in ids]], 64)
    assert np.array_equal(km.read(), want)
    kb = r["kmer_by_matrix_column"].read()
    assert kb.dtype == np.uint16 and np.array_equal(kb, np.arange(U))


def test_from_tsv_without_phenotype_keeps_file_order(tmp_path):
    rng = np.random.default_rng(12)
    kmers, present, mat = _random_matrix(rng, 5, 40, 9)
    names = list("edcba")
    p = tmp_path / "m.tsv"
    tsv.write_tsv(str(p), kmers, mat, names, 9)
    out = tmp_path / "o.kover"
    create.from_tsv(str(p), str(out), None, None, 0)
    r = hdf5min.H5Reader(str(out))
    assert [x.decode() for x in r["genome_identifiers"].read()] == names
    assert r.attrs["phenotype_description"] == "NA" and "phenotype" not in r and "classification_type" not in r.attrs
    assert np.array_equal(r["kmer_matrix"].read(), mat)
    with pytest.raises(ValueError):
        create.from_tsv(str(p), str(out), "desc only", None, 0)


def test_tsv_reader_enforces_fixed_width_rows(tmp_path):
    p = tmp_path / "bad.tsv"
    p.write_bytes(b"kmers\ta\tb\nACG\t1\t0\nACGT\t1\t1\n")
    with pytest.raises(Exception):
        create.read_kmer_matrix_tsv(str(p))
    p.write_bytes(b"kmers\ta\tb\n")
    ids, km, cells = create.read_kmer_matrix_tsv(str(p))
    assert ids == ["a", "b"] and len(km) == 0 and cells.shape == (0, 2)


# ---- CLI surface (bin/kover/kover:36-224) ---------------------------------------------------------------
def test_cli_from_tsv_and_errors(tmp_path, capsys):
    rng = np.random.default_rng(13)
    kmers, _, mat = _random_matrix(rng, 4, 25, 11)
    p = tmp_path / "m.tsv"
    tsv.write_tsv(str(p), kmers, mat, ["a", "b", "c", "d"], 11)
    out = tmp_path / "o.kover"
    assert cli.main(["dataset", "create", "from-tsv", "--genomic-data", str(p), "--output", str(out)]) == 0
    assert hdf5min.H5Reader(str(out)).attrs["compression"] == "gzip (level 4)"      # kover:119 default
    assert cli.main(["dataset", "create", "from-tsv", "--genomic-data", str(p), "--output", str(out),
                     "--phenotype-description", "x"]) == 1
    assert cli.main(["dataset", "split"]) == 2
    assert cli.main(["--version"]) == 0
    # a command built by the GUI-side builder parses with this CLI
    cmd = kover_cmd.create_command("kover", kover_cmd.Source.K_MER_MATREX, str(p), str(out), compression=2, x=True)
    assert cli.main(cmd.split()[1:]) == 0
    assert hdf5min.H5Reader(str(out)).attrs["compression"] == "gzip (level 2)"
    # missing input -> non-zero exit, message on stderr (reference prints and exits, kover:133-136)
    assert cli.main(["dataset", "create", "from-tsv", "--genomic-data", str(tmp_path / "nope"), "--output", str(out)]) == 1
    assert "Error" in capsys.readouterr().err


def test_build_pipeline_slots_and_order(monkeypatch):
    """BuildPipeline's host logic without a GPU: submissions alternate over the slots, every slot works through its
    own submissions in order (one worker each), results come back through the futures, close() closes every slot."""
    import threading
    import time
    from grm_b200 import builder as B

    class FakeBuilder:
        made = []

        def __init__(self, **kw):
            self.kw, self.log, self.closed = kw, [], False
            self._n = 0
            FakeBuilder.made.append(self)

        def reset(self):
            self.log.append("reset")

        def set_genome_count(self, n):
            self.log.append(("count", n))

        def add_genomes(self, rows, data, lens, on_device):
            self.log.append(("add", list(rows), on_device))

        def build(self):
            time.sleep(0.01)
            self._n += 1

        def result_host(self):
            return ("km", self._n, threading.get_ident()), ("mat", self._n)

        @property
        def stats(self):
            return {"n": self._n}

        def close(self):
            self.closed = True

    monkeypatch.setattr(B, "KmerMatrixBuilder", FakeBuilder)
    with B.BuildPipeline(depth=2, streams=[11, 22], k=31, keep_singletons=True) as pipe:
        slots = list(pipe.slots)
        assert [s.kw["stream"] for s in slots] == [11, 22] and all(s.kw["k"] == 31 for s in slots)
        futs = [pipe.submit([i], [b"x"], n_genomes=i) for i in range(5)]
        res = [f.result() for f in futs]
    # submissions 0, 2, 4 went to slot 0 (its 1st, 2nd, 3rd build), 1 and 3 to slot 1
    assert [r[0][1] for r in res] == [1, 1, 2, 2, 3]
    assert len({res[0][0][2], res[2][0][2], res[4][0][2]}) == 1 and res[0][0][2] != res[1][0][2]      # one worker thread per slot
    assert [e for e in slots[0].log if e[0] == "add"] == [("add", [0], False), ("add", [2], False), ("add", [4], False)]
    assert ("count", 3) in slots[1].log and ("count", 0) not in slots[0].log            # n_genomes = 0 means "not given"
    assert all(s.closed for s in slots)
    with pytest.raises(ValueError):
        B.BuildPipeline(depth=0)
    with pytest.raises(ValueError):
        B.BuildPipeline(depth=2, streams=[1])
