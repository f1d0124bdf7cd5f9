"""Owner-side merge of an N-GPU build, timed on ONE GPU (no collective is needed to study the kernel).

    python tools/probe/merge_bench.py [world=8] [genomes_per_rank=125] [reps=5] [owner=0] [--check]

The `world` emulated ranks build their partial columns one after the other on this GPU (same synthetic genomes, same
64-aligned row blocks and hash-range owners as bench.py's C3 family), the owner's receive buffer is assembled by
slicing, and grmkm_merge_partials is timed by its own stage events.  With --check the owner's slice checksum is
printed next to the checksum of a one-GPU build restricted to the owner's hash range (first / last k-mer of the slice).
"""
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))


def main():
    import torch
    import bench
    from grm_b200 import synth
    from grm_b200.builder import KmerMatrixBuilder
    from grm_b200.distributed import CudaEngine, row_partition, words_per_rank
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    world = int(args[0]) if len(args) > 0 else 8
    per = int(args[1]) if len(args) > 1 else 125
    reps = int(args[2]) if len(args) > 2 else 5
    owner = int(args[3]) if len(args) > 3 else 0
    G = world * per
    cfg = synth.SynthConfig(seed=synth.MASTER_SEED + 1)
    parts, sw = row_partition(G, world), words_per_rank(G, world)
    b = KmerMatrixBuilder(k=31, keep_singletons=True)
    e = CudaEngine(b)
    chunks, src_counts, bits = [], [], 0
    # bucket bits: the maximum of the ranks' plans (here: every rank has the same amount of text)
    for r in range(world):
        buf, spans, _ = bench.device_genomes(b, cfg, parts[r])
        b.reset()
        b.set_genome_count(len(parts[r]))
        b.add_genomes(np.arange(len(spans), dtype=np.uint32), np.array([buf.data_ptr() + o for o, _ in spans], dtype=np.uint64),
                      np.array([n for _, n in spans], dtype=np.uint64), on_device=True)
        if r == 0:
            bits = e.plan_bucket_bits()
        e.set_bucket_bits(bits)
        counts, send = e.build_partial(world, sw[r])
        width = 1 + sw[r]
        off = sum(counts[:owner]) * width
        chunks.append(send[off: off + counts[owner] * width].clone())
        src_counts.append(counts[owner])
        print(f"rank {r}: {sum(counts)} partial columns, {counts[owner]} for owner {owner}, local stages {e.local_times}", flush=True)
        del buf, send
    recv = torch.cat(chunks)
    del chunks
    torch.cuda.synchronize()
    ts = []
    for i in range(reps):
        t0 = time.perf_counter()
        e.merge(recv, world, owner, src_counts, sw, G)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        ts.append((dt, dict(e.merge_times)))
    for dt, mt in ts:
        print("merge wall %.3f ms  stages %s" % (dt, {k: round(v, 3) for k, v in mt.items() if v}))
    U = b.dims[0]
    print(f"owner {owner}: {sum(src_counts)} entries in ({recv.numel() * 8 / 1e6:.1f} MB), {U} columns out "
          f"({U * (8 + 8 * ((G + 63) // 64)) / 1e6:.1f} MB), checksum {b.checksum()}")
    b.close()


if __name__ == "__main__":
    main()
