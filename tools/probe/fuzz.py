"""Randomised parity sweep (GPU vs oracle) over input shapes and build parameters.  usage: fuzz.py [iterations] [seed]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import oracle
from tests import inputs
from grm_b200.builder import KmerMatrixBuilder

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 12345
rng = np.random.default_rng(seed)
t0 = time.time()
for it in range(iters):
    G = int(rng.choice([1, 2, 3, 7, 20, 64, 65, 130, 260]))
    k = int(rng.choice([5, 9, 11, 15, 16, 17, 21, 25, 31, 32]))
    kind = int(rng.random() < 0.25)
    keep = bool(rng.random() < 0.5)
    min_ab = int(rng.choice([1, 1, 1, 2, 3])) if kind else 1
    bucket_bits = int(rng.choice([0, 0, 0, 6, 9, 12, 14]))
    big = rng.random() < 0.2
    shared = [inputs.rand_seq(rng, int(rng.integers(200, 60000 if big else 4000))) for _ in range(int(rng.integers(1, 4)))]
    genomes = []
    for g in range(G):
        files = []
        for _ in range(int(rng.choice([1, 1, 1, 2]))):
            if kind:
                files.append(inputs.fastq(rng, shared[int(rng.integers(len(shared)))], n_reads=int(rng.integers(0, 400 if big else 60)),
                                          read_len=int(rng.integers(10, 300)), crlf=bool(rng.random() < 0.2),
                                          final_nl=bool(rng.random() < 0.8)))
            else:
                files.append(inputs.fasta(rng, n_records=int(rng.integers(0, 6)), max_len=int(rng.integers(1, 30000 if big else 900)),
                                          width=int(rng.choice([0, 7, 60, 61, 80, 4099])), crlf=bool(rng.random() < 0.2),
                                          blank=bool(rng.random() < 0.2), final_nl=bool(rng.random() < 0.8),
                                          junk_prefix=bool(rng.random() < 0.1), shared=shared,
                                          p_n=float(rng.choice([0.0, 0.01, 0.2])), p_lower=float(rng.choice([0.0, 0.05, 1.0]))))
        genomes.append(files)
    ref = oracle.build([[(f, kind) for f in files] for files in genomes], k, min_ab, keep)
    with KmerMatrixBuilder(k=k, min_abundance=min_ab, keep_singletons=keep, input_kind=kind, bucket_bits=bucket_bits) as b:
        b.set_genome_count(G)
        rows = [r for r, files in enumerate(genomes) for _ in files]
        b.add_genomes(rows, [f for files in genomes for f in files])
        b.build()
        km, mat, st = b.kmers(), b.matrix(), b.stats
    ok = (np.array_equal(km, ref.kmers) and np.array_equal(mat, ref.matrix) and st["n_bases"] == ref.n_bases
          and st["n_windows"] == ref.n_windows)
    if not ok:
        print(f"MISMATCH it={it} seed={seed} G={G} k={k} kind={kind} keep={keep} min_ab={min_ab} bits={bucket_bits} big={big} "
              f"U={len(km)} vs {len(ref.kmers)} stats={st}", flush=True)
        sys.exit(1)
print(f"fuzz ok: {iters} builds, {time.time() - t0:.1f} s, seed {seed}")
