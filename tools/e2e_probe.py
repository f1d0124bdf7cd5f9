"""Where the end-to-end (host buffers in, host arrays out) time goes."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from grm_b200 import synth  # noqa: E402
from grm_b200.builder import KmerMatrixBuilder  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 100
cfg = synth.SynthConfig(seed=synth.MASTER_SEED + 1)
lay, total, spans = synth.build_layout(cfg, range(G))
buf = torch.empty(total, dtype=torch.uint8, device="cuda")
with KmerMatrixBuilder(k=31, keep_singletons=True) as b:
    b._check(b._lib.grmkm_synth_fasta_device(b._ctx, C.c_void_p(lay.ctypes.data), lay.nbytes, C.c_void_p(buf.data_ptr()), total))
    host = [torch.empty(ln, dtype=torch.uint8).pin_memory() for _, ln in spans]
    for t, (off, ln) in zip(host, spans):
        t.copy_(buf[off:off + ln])
    torch.cuda.synchronize()
    arrs = [t.numpy() for t in host]
    for rep in range(4):
        t0 = time.perf_counter()
        b.reset()
        for i, a in enumerate(arrs):
            b.add_genome_bytes(i, a)
        t1 = time.perf_counter()
        b.build()
        t2 = time.perf_counter()
        km = b.kmers()
        t3 = time.perf_counter()
        mat = b.matrix()
        t4 = time.perf_counter()
        print(f"rep{rep}: add {1e3*(t1-t0):.2f}  build {1e3*(t2-t1):.2f} (dev total {b.times['total']:.2f}, h2d {b.times['h2d']:.2f})  "
              f"kmers {1e3*(t3-t2):.2f}  matrix {1e3*(t4-t3):.2f}  sum {1e3*(t4-t0):.2f} ms")
