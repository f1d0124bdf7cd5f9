import sys
import numpy as np
sys.path.insert(0, ".")
from tests import inputs
from grm_b200 import native
from grm_b200.builder import KmerMatrixBuilder
rng = np.random.default_rng(1)
genomes = [inputs.fasta(rng, n_records=3, max_len=3000) for _ in range(5)]
for flags in (0, native.FLAG_KMER_RECORDS, native.FLAG_KMER_RECORDS, 0, native.FLAG_KMER_RECORDS):
    with KmerMatrixBuilder(k=31, keep_singletons=True, flags=flags) as b:
        for r, g in enumerate(genomes):
            b.add_genome_bytes(r, g)
        try:
            b.build()
            print("flags", flags, "ok", b.dims, {k: round(v, 3) for k, v in b.times.items()})
        except Exception as e:
            print("flags", flags, "FAIL", e)
