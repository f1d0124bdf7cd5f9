"""Ad-hoc timing probe (not the bench): synthetic genomes generated on the device, one build, stage times."""
import ctypes as C
import sys
import time

import torch

sys.path.insert(0, ".")
from grm_b200 import synth  # noqa: E402
from grm_b200.builder import KmerMatrixBuilder  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
bbits = int(sys.argv[4]) if len(sys.argv) > 4 else 0
flags = int(sys.argv[5]) if len(sys.argv) > 5 else 0
cfg = synth.SynthConfig(seed=synth.MASTER_SEED + 1)
if scale != 1.0:
    cfg = cfg.scaled(scale)
t0 = time.time()
lay, total, spans = synth.build_layout(cfg, range(G))
print(f"layout {time.time()-t0:.2f}s total_bytes={total}")
buf = torch.empty(total, dtype=torch.uint8, device="cuda")
with KmerMatrixBuilder(k=31, keep_singletons=True, bucket_bits=bbits, flags=flags) as b:
    b._check(b._lib.grmkm_synth_fasta_device(b._ctx, C.c_void_p(lay.ctypes.data), lay.nbytes, C.c_void_p(buf.data_ptr()), total))
    for r in range(reps):
        b.reset()
        for row, (off, ln) in enumerate(spans):
            b.add_genome_device(row, buf.data_ptr() + off, ln)
        t0 = time.time()
        b.build()
        dt = time.time() - t0
        s, t = b.stats, b.times
        print(f"rep{r}: wall {dt*1e3:.2f} ms  dev {t['total']:.3f} ms  Gbases/s {s['n_bases']/t['total']/1e6:.2f}")
        print("   times", {k: round(v, 3) for k, v in t.items()})
    print("   stats", s)
