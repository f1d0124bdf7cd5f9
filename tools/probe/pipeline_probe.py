"""Timeline of BuildPipeline on the C2 workload: when every build starts, returns from grmkm_build and has its
result on the host (wall clock, ms from the first submission).  usage: pipeline_probe.py [genomes=100] [builds=8] [depth=2]"""
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))


def main():
    import torch
    import bench
    from grm_b200 import synth
    from grm_b200.builder import BuildPipeline, KmerMatrixBuilder
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    depth = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    cfg = synth.SynthConfig(seed=synth.MASTER_SEED + 1)
    with KmerMatrixBuilder(k=31, keep_singletons=True) as b0:
        buf, spans, _ = bench.device_genomes(b0, cfg, range(G))
    offs, tot = [], 0
    for _, ln in spans:
        offs.append(tot)
        tot += (ln + 15) & ~15
    host = torch.zeros(tot, dtype=torch.uint8).pin_memory()          # one arena, files 16-byte aligned (create._read_inputs)
    for o, (off, ln) in zip(offs, spans):
        host[o:o + ln].copy_(buf[off:off + ln])
    torch.cuda.synchronize()
    arena = host.numpy()
    host_np = [arena[o:o + ln] for o, (_, ln) in zip(offs, spans)]
    rows = np.arange(G, dtype=np.uint32)
    log = []
    t_ref = [0.0]

    def run(b, rows_, data, lens, on_device, n_genomes):
        i = len(log)
        rec = {"i": i, "start": time.perf_counter()}
        log.append(rec)
        b.reset()
        b.add_genomes(rows_, data, lens, on_device)
        rec["added"] = time.perf_counter()
        b.build()
        rec["built"] = time.perf_counter()
        km, mat = b.result_host()
        rec["host"] = time.perf_counter()
        rec["times"] = b.times
        return km, mat, None

    BuildPipeline._run = staticmethod(run)
    with BuildPipeline(depth=depth, k=31, keep_singletons=True) as pipe:
        for f in [pipe.submit(rows, host_np) for _ in range(2 * depth)]:
            f.result()
        log.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for f in [pipe.submit(rows, host_np) for _ in range(n)]:
            f.result()
        t1 = time.perf_counter()
    print(f"{n} builds of {G} genomes, depth {depth}: {(t1 - t0) * 1e3 / n:.2f} ms per build")
    for r in sorted(log, key=lambda r: r["i"]):
        print("build %2d  start %7.2f  added %7.2f  built %7.2f  on host %7.2f   (build %.2f, d2h %.2f; kernels total %.2f, scatter stage %.2f)" % (
            r["i"], (r["start"] - t0) * 1e3, (r["added"] - t0) * 1e3, (r["built"] - t0) * 1e3, (r["host"] - t0) * 1e3,
            (r["built"] - r["added"]) * 1e3, (r["host"] - r["built"]) * 1e3, r["times"]["total"], r["times"]["scatter"]))


if __name__ == "__main__":
    main()
