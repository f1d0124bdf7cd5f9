"""GPU: the reference-facing entry points end to end (files in -> .kover HDF5 / KmerMatrix.tsv out)
against the CPU oracle: from_contigs / from_reads (create.py:278-523), the kover-compatible CLI,
the tool shims (tools/kmer_count.py, kmer_pack.py) and the Ray Surveyor runner (app.py:1280-1354)."""
import gzip as gz
import os

import numpy as np
import pytest

from oracle import oracle
from tests import inputs

pytestmark = pytest.mark.gpu


def _write_genomes(tmp_path, rng, G, k, reads=False):
    shared = [inputs.rand_seq(rng, 1500), inputs.rand_seq(rng, 600)]
    ids, paths, texts = [], [], []
    for g in range(G):
        gid = f"562.{g}"
        if reads:
            d = tmp_path / f"reads_{g}"
            d.mkdir()
            src = shared[0] if g % 3 else shared[0][:900] + inputs.rand_seq(rng, 300)
            a = inputs.fastq(rng, src, n_reads=150, read_len=60, err=0.01)
            b = inputs.fastq(rng, src, n_reads=100, read_len=60, err=0.01)
            (d / "r_1.fastq").write_bytes(a)
            with gz.open(d / "r_2.fastq.gz", "wb") as f:
                f.write(b)
            (d / "notes.txt").write_text("ignored")
            ids.append(gid); paths.append(str(d)); texts.append([a, b])
        else:
            fa = inputs.fasta(rng, n_records=4, shared=shared, max_len=500, crlf=(g == 2), blank=(g == 3))
            p = tmp_path / f"{gid}.fna"
            p.write_bytes(fa)
            ids.append(gid); paths.append(str(p)); texts.append([fa])
    return ids, paths, texts


def _expect(ids_in_file_order, texts, order_ids, k, m, keep, kind):
    row = {g: i for i, g in enumerate(ids_in_file_order)}
    genomes = [[(t, kind) for t in texts[row[g]]] for g in order_ids]
    return oracle.build(genomes, k, m, keep)


def _check_kover(path, ref, k, filter_name, source):
    from grm_b200 import hdf5min
    r = hdf5min.H5Reader(str(path))
    assert r.attrs["genome_source_type"] == source and r.attrs["filter"] == filter_name
    seqs = r["kmer_sequences"].read()
    assert [s.decode() for s in seqs] == [oracle.py_kmer_string(int(x), k) for x in ref.kmers]
    km = r["kmer_matrix"]
    assert km.chunks == (1, max(1, min(ref.n_kmers, 100000))) or ref.n_kmers == 0
    assert np.array_equal(km.read(), ref.matrix)
    assert np.array_equal(r["kmer_by_matrix_column"].read(), np.arange(ref.n_kmers))
    return r


@pytest.mark.parametrize("singleton", [False, True])
def test_from_contigs_end_to_end(gpu, tmp_path, singleton):
    from grm_b200 import create
    rng = np.random.default_rng(21)
    k = 21
    ids, paths, texts = _write_genomes(tmp_path, rng, 9, k)
    lst = tmp_path / "genomes_paths.tsv"
    lst.write_text("".join(f"{g}\t{p}\n" for g, p in zip(ids, paths)))
    meta = tmp_path / "meta.tsv"
    lab = ["S", "R", "R", "S", "R", "S", "S", "R"]                 # genome 8 has no metadata -> dropped
    meta.write_text("".join(f"{g}\t{l}\n" for g, l in zip(ids, lab)))
    out = tmp_path / "DATASET.kover"
    filt = "nothing" if singleton else "singleton"
    warnings = []
    create.from_contigs(str(lst), str(out), k, filt, "desc", str(meta), 4, str(tmp_path), 2, False, False,
                        warnings.append)
    assert len(warnings) == 1
    from grm_b200 import hdf5min
    got_ids = [x.decode() for x in hdf5min.H5Reader(str(out))["genome_identifiers"].read()]
    assert sorted(got_ids) == sorted(ids[:8])
    ref = _expect(ids, texts, got_ids, k, 1, singleton, 0)
    r = _check_kover(out, ref, k, filt, "contigs")
    ph = r["phenotype"].read().tolist()
    assert ph == sorted(ph) and [lab[ids.index(g)] for g in got_ids] == ["RS"[p] for p in ph]


def test_from_contigs_without_phenotype_and_missing_file(gpu, tmp_path):
    from grm_b200 import create
    rng = np.random.default_rng(22)
    ids, paths, texts = _write_genomes(tmp_path, rng, 3, 15)
    lst = tmp_path / "l.tsv"
    lst.write_text("".join(f"{g} {p}\n" for g, p in zip(ids, paths)))
    out = tmp_path / "o.kover"
    create.from_contigs(str(lst), str(out), "15", "singleton", None, None, 0, None, 0, False, False)
    ref = _expect(ids, texts, ids, 15, 1, False, 0)
    _check_kover(out, ref, 15, "singleton", "contigs")
    lst.write_text(f"{ids[0]} {paths[0]}\nghost {tmp_path}/nope.fna\n")
    with pytest.raises(IOError):
        create.from_contigs(str(lst), str(out), 15, "singleton", None, None, 0, None, 0, False, False)


@pytest.mark.parametrize("abundance", [1, 2])
def test_from_reads_end_to_end(gpu, tmp_path, abundance):
    from grm_b200 import cli
    rng = np.random.default_rng(23)
    k = 15
    ids, dirs, texts = _write_genomes(tmp_path, rng, 5, k, reads=True)
    lst = tmp_path / "reads.tsv"
    lst.write_text("".join(f"{g}\t{d}\n" for g, d in zip(ids, dirs)))
    out = tmp_path / "reads.kover"
    rc = cli.main(["dataset", "create", "from-reads", "--genomic-data", str(lst), "--output", str(out),
                   "--kmer-size", str(k), "--kmer-min-abundance", str(abundance), "--singleton-kmers",
                   "--compression", "1", "--n-cpu", "2"])
    assert rc == 0
    # file order inside a genome's directory is os.listdir order; pooling makes it irrelevant
    ref = _expect(ids, texts, ids, k, abundance, True, 1)
    _check_kover(out, ref, k, "nothing", "reads")


def test_cli_reports_failure_with_nonzero_exit(gpu, tmp_path):
    from grm_b200 import cli
    lst = tmp_path / "l.tsv"
    lst.write_text(f"g1 {tmp_path}/missing.fna\n")
    assert cli.main(["dataset", "create", "from-contigs", "--genomic-data", str(lst), "--output",
                     str(tmp_path / "o.kover")]) == 1
    p = tmp_path / "a.fna"
    p.write_bytes(b">a\nACGTACGTAC\n")
    lst.write_text(f"g1 {p}\n")
    assert cli.main(["dataset", "create", "from-contigs", "--genomic-data", str(lst), "--output",
                     str(tmp_path / "o.kover"), "--kmer-size", "33"]) == 1            # k > 32 unsupported


def test_tool_shims_count_then_pack(gpu, tmp_path):
    """kmer_count.py:23-37 then kmer_pack.py:23-36 with the reference's argument lists."""
    from grm_b200 import hdf5min, tools
    rng = np.random.default_rng(24)
    k = 11
    ids, paths, texts = _write_genomes(tmp_path, rng, 4, k)
    tmp = tmp_path / "tmp"
    tmp.mkdir()
    (tmp / "list_contigs_files").write_text("\n".join(paths))
    out = tmp_path / "d.kover"
    with hdf5min.H5Writer(str(out)) as h5:                       # the skeleton create.py:311-356 writes first
        h5.attrs["filter"] = "singleton"
        h5.create_dataset("genome_identifiers", np.array(ids).astype("S"))
    tools.contigs_count_kmers(str(tmp / "list_contigs_files"), str(tmp), k, 4, 2, False, False)
    (tmp / "list_h5").write_text("\n".join(str(tmp / (os.path.basename(p)[:-4] + ".h5")) for p in paths))
    tools.contigs_pack_kmers(str(tmp / "list_h5"), str(out), "singleton", k, 4, 100000, len(ids), False)
    ref = _expect(ids, texts, ids, k, 1, False, 0)
    r = hdf5min.H5Reader(str(out))
    assert r.attrs["filter"] == "singleton" and [x.decode() for x in r["genome_identifiers"].read()] == ids
    assert np.array_equal(r["kmer_matrix"].read(), ref.matrix)
    assert [s.decode() for s in r["kmer_sequences"].read()] == [oracle.py_kmer_string(int(x), k) for x in ref.kmers]


def test_surveyor_conf_to_tsv_to_kover(gpu, tmp_path):
    """survey.conf -> KmerMatrix.tsv (Ray stand-in) -> from_tsv: the GUI's two-step route (Readme.md:82-84)."""
    from grm_b200 import create, hdf5min, surveyor
    rng = np.random.default_rng(25)
    k = 31
    ids, paths, texts = _write_genomes(tmp_path, rng, 6, k)
    out_dir = tmp_path / "survey out"
    out_dir.mkdir()
    conf = surveyor.generate_survey_conf(paths, k, str(out_dir))
    assert surveyor.main([conf]) == 0
    tsv_path = out_dir / "survey.res" / "Surveyor" / "KmerMatrix.tsv"
    names = [os.path.basename(p)[:-4] for p in paths]
    ref = oracle.build([[(t, 0) for t in ts] for ts in texts], k, 1, True, want_tsv_names=names)
    assert tsv_path.read_bytes() == ref.tsv
    out = tmp_path / "from_tsv.kover"
    create.from_tsv(str(tsv_path), str(out), None, None, 4)
    r = hdf5min.H5Reader(str(out))
    assert np.array_equal(r["kmer_matrix"].read(), ref.matrix)
    assert surveyor.main([str(tmp_path / "missing.conf")]) == 1


def test_build_pipeline_series_of_datasets(gpu):
    """BuildPipeline: several builds in flight on one GPU (two contexts, one worker thread each, H2D legs gated one
    after the other by the library).  Every dataset of the series equals the oracle's, whatever slot built it, and the
    slots' result buffers stay valid for `depth` submissions."""
    from grm_b200.builder import BuildPipeline
    from grm_b200.synth import SynthConfig, genome_fasta, MASTER_SEED
    rng = np.random.default_rng(77)
    cfg = SynthConfig(seed=MASTER_SEED + 5).scaled(0.02)
    sets = []
    for d in range(7):
        G = int(rng.integers(1, 9))
        texts = [genome_fasta(cfg, int(g)) for g in rng.integers(0, 40, size=G)]
        sets.append(texts)
    with BuildPipeline(depth=2, k=31, keep_singletons=True) as pipe:
        futs = [pipe.submit(np.arange(len(t), dtype=np.uint32), t, n_genomes=len(t)) for t in sets]
        got = []
        for f in futs:
            km, mat, st = f.result()
            got.append((km.copy(), mat.copy(), st))       # the views die two submissions later
    for texts, (km, mat, st) in zip(sets, got):
        ref = oracle.build([[(t, 0)] for t in texts], 31, 1, True)
        assert np.array_equal(km, ref.kmers) and np.array_equal(mat, ref.matrix)
        assert st["n_windows"] == ref.n_windows
