import sys
sys.path.insert(0, ".")
import numpy as np
from oracle import oracle
from tests import inputs
from tests.test_gpu_parity import gpu_build

rng = np.random.default_rng(51)
src = [inputs.rand_seq(rng, 3000) for _ in range(3)]
files = []
for g in range(6):
    s = src[g % 3]
    files.append(inputs.fastq(rng, s, n_reads=400, read_len=60, crlf=(g == 1), final_nl=(g != 2)))
    files.append(inputs.fastq(rng, s, n_reads=100, read_len=45))
for i, f in enumerate(files):
    ref = oracle.build([[(f, 1)]], 21, 1, True)
    km, mat, st = gpu_build([[f]], 21, 1, True, 1)
    print(i, len(f), "ref windows", ref.n_windows, "gpu", st["n_windows"], "bases", ref.n_bases, st["n_bases"], "U", ref.n_kmers, st["n_kmers"])
# shrink: first failing file, try prefixes
for i, f in enumerate(files):
    ref = oracle.build([[(f, 1)]], 21, 1, True)
    km, mat, st = gpu_build([[f]], 21, 1, True, 1)
    if ref.n_windows != st["n_windows"]:
        lines = f.split(b"\n")
        for nrec in [1, 2, 3, 5, 10, 20, 40, 80]:
            sub = b"\n".join(lines[:4 * nrec]) + b"\n"
            ref = oracle.build([[(sub, 1)]], 21, 1, True)
            km, mat, st = gpu_build([[sub]], 21, 1, True, 1)
            print("  file", i, "nrec", nrec, len(sub), ref.n_windows, st["n_windows"])
            if ref.n_windows != st["n_windows"] and nrec <= 3:
                print(sub)
        break
