#!/usr/bin/env python
"""bench.py -- Gbases/s of the k-mer matrix build (k=31) on N B200s, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W]          (N>1: launched by torch.distributed.run)
    python bench.py --impl reference [...]                        CPU arm: the oracle port on the host cores

A "step" is one complete build of the workload: FASTA text -> packed stream -> minimizer-bounded units
(super-k-mers) -> content-hash buckets -> per-bucket dedupe across genomes -> k-mers of the distinct units ->
hash buckets -> shared-memory aggregation -> ordered columns (kmers[U] + matrix[W][U]).
  value : whole-job Gbases/s with the FASTA bytes already resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through the public API with HOST (pinned) buffers: H2D of the text and D2H of
          kmers + matrix inside the timed region
Workload at N=1 = BASELINE.json configs[1] (Ray Surveyor matrix: 100 synthetic 5 Mbp genomes, k=31,
min abundance 1, no singleton filter); at N>1 = configs[2] family (125 genomes per GPU, 1000 at N=8,
rows split across ranks in 64-aligned blocks, partial columns exchanged by hash range: the export kernel stores
every owner's slice into its receive buffer over NVLink (torch symmetric memory; NCCL all-to-all as the fallback)).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gbases/s k-mer matrix build (k=31)"
UNIT = "Gbases/s"
K = 31
GENOMES_N1 = 100          # configs[1]
GENOMES_PER_GPU = 125     # configs[2]: 1000 genomes on 8 GPUs
CPU_SAMPLE_GENOMES = 100


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_from_profiles(kernel: str):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            v = json.load(f).get(kernel)
            return v.get("dram_bytes_per_launch") if isinstance(v, dict) else v
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_arm(genome_texts, threads=0):
    """The oracle port (restatement of multidsk + dsk2kover) on the host cores.  Checker-side code:
    only this baseline leg and --impl reference may execute oracle/."""
    from oracle import oracle
    oracle.build_library()
    t0 = time.perf_counter()
    r = oracle.build([[(t, 0)] for t in genome_texts], K, 1, True, threads=threads)
    dt = time.perf_counter() - t0
    return r.n_bases / dt / 1e9, dt, oracle.threads() if threads <= 0 else threads, r


def device_genomes(builder, cfg, ids):
    """Synthesise the genomes' FASTA on the device; returns (uint8 cuda tensor, spans, n_bases)."""
    import torch
    from grm_b200 import synth
    lay, total, spans = synth.build_layout(cfg, ids)
    buf = torch.empty(max(total, 16), dtype=torch.uint8, device="cuda")
    builder._check(builder._lib.grmkm_synth_fasta_device(builder._ctx, C.c_void_p(lay.ctypes.data), lay.nbytes,
                                                          C.c_void_p(buf.data_ptr()), total))
    return buf, spans, synth.n_bases_of(cfg, ids)


def run_reference(args):
    """--impl reference: the CPU implementation of the path (oracle port: the reference binaries are
    absent from the checkout, .MISSING_LARGE_BLOBS:1-4) on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from grm_b200 import synth
    cfg = synth.SynthConfig(seed=synth.MASTER_SEED + 1)
    n_gen = CPU_SAMPLE_GENOMES
    texts = None
    try:
        import torch
        if torch.cuda.is_available():
            from grm_b200.builder import KmerMatrixBuilder
            with KmerMatrixBuilder(k=K, keep_singletons=True) as b:
                buf, spans, _ = device_genomes(b, cfg, range(n_gen))
                host = buf.cpu().numpy()
                texts = [host[o:o + n].tobytes() for o, n in spans]
    except Exception:
        texts = None
    if texts is None:
        texts = [synth.genome_fasta(cfg, g) for g in range(n_gen)]
    vals, times = [], []
    for i in range(args.warmup + args.steps):
        v, dt, threads, _ = cpu_arm(texts)
        if i >= args.warmup:
            vals.append(v); times.append(dt)
    val = sum(vals) / len(vals)
    sample = f"{n_gen} of the workload's synthetic 5 Mbp genomes per step (k=31, min abundance 1, singletons kept)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(n_gpus):
    if n_gpus == 1:
        wl = f"C2: Ray Surveyor genome x k-mer matrix, {GENOMES_N1} synthetic 5 Mbp genomes, k=31, min abundance 1, singletons kept"
        g = GENOMES_N1
    else:
        g = GENOMES_PER_GPU * n_gpus
        wl = (f"C3 family: {g} synthetic 5 Mbp genomes ({GENOMES_PER_GPU} per GPU) k=31, rows sharded across "
              f"{n_gpus} GPUs in 64-aligned blocks, hash-range exchange (NVLink peer stores fused into the export kernel)")
    return {"workload": wl, "genomes": g, "k": K, "min_abundance": 1, "keep_singletons": True,
            "l2": "inputs (>=500 MB FASTA per GPU) larger than the 126 MB L2; no flush needed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="grm_b200", choices=["grm_b200", "reference"])
    ap.add_argument("--genomes", type=int, default=0, help="override genomes per GPU (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    from grm_b200 import synth
    from grm_b200.builder import KmerMatrixBuilder
    from grm_b200.distributed import DistributedBuilder, init_process_group_from_env

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the k-mer matrix path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = init_process_group_from_env() if world > 1 else None
    if args.warmup < 3 and rank == 0:
        print("note: fewer than 3 warm-up steps; this run is for profiling, not a bench value", file=sys.stderr)

    cfg = synth.SynthConfig(seed=synth.MASTER_SEED + 1)
    if world == 1:
        G_total = args.genomes or GENOMES_N1
    else:
        G_total = (args.genomes or GENOMES_PER_GPU) * world
    stream = torch.cuda.Stream()
    peak, peak_src = peaks()

    with torch.cuda.stream(stream):
        db = DistributedBuilder(k=K, min_abundance=1, keep_singletons=True, n_genomes=G_total, rank=rank, world=world,
                                stream=stream.cuda_stream, device=local_rank)
        my_rows = list(db.local_rows)                     # global genome rows of this rank
        buf, spans, my_bases = device_genomes(db.builder, cfg, my_rows)
        stream.synchronize()

        rows_np = np.arange(len(spans), dtype=np.uint32)
        ptrs_np = np.array([buf.data_ptr() + off for off, _ in spans], dtype=np.uint64)
        lens_np = np.array([ln for _, ln in spans], dtype=np.uint64)

        def step_resident():
            db.reset()
            db.add_genomes(rows_np, ptrs_np, lens_np, on_device=True)
            db.build(reuse_partition=True)        # N > 1: the bucket count agreed in the first warm-up build is kept

        # ---- value: inputs resident in HBM
        for _ in range(args.warmup):
            step_resident()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        if dist:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage_acc, launches = {}, 0
        e0.record(stream)
        for _ in range(args.steps):
            step_resident()
            for kname, v in db.stage_times.items():
                stage_acc[kname] = stage_acc.get(kname, 0.0) + v
            launches += db.launches
        e1.record(stream)
        torch.cuda.synchronize()
        if dist:
            torch.distributed.barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
        stats = dict(db.builder.stats)
        stats.update({k2: v for k2, v in db.local_stats.items() if k2 in ("n_windows", "n_input_bytes", "n_bases", "n_records", "n_buckets", "n_units", "n_unit_entries", "n_wide", "n_unit_buckets")})
        U_local = db.n_kmers

        # ---- e2e: host (pinned) buffers in, kmers + matrix out, every step
        host_in = [torch.empty(ln, dtype=torch.uint8).pin_memory() for _, ln in spans]
        for t, (off, ln) in zip(host_in, spans):
            t.copy_(buf[off:off + ln])
        stream.synchronize()
        host_np = [t.numpy() for t in host_in]
        h2d = sum(ln for _, ln in spans)

        def step_e2e():
            db.reset()
            db.add_genomes(rows_np, host_np)
            db.build(reuse_partition=True)
            return db.result_host()          # one D2H copy into the context's page-locked result buffer

        for _ in range(min(args.warmup, 2)):
            km, mat = step_e2e()
        if dist:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2e_steps = max(1, min(args.steps, 5))
        f0.record(stream)
        for _ in range(e2e_steps):
            km, mat = step_e2e()
        f1.record(stream)
        torch.cuda.synchronize()
        if dist:
            torch.distributed.barrier()
        ms_e2e = f0.elapsed_time(f1)
        d2h = int(km.nbytes + mat.nbytes)

    # ---- reduce over ranks: max time, sum bases
    tt = torch.tensor([ms, ms_e2e, float(my_bases), float(U_local), float(launches), float(h2d), float(d2h)],
                      dtype=torch.float64, device="cuda")
    if dist:
        mx = tt.clone(); torch.distributed.all_reduce(mx, op=torch.distributed.ReduceOp.MAX)
        sm = tt.clone(); torch.distributed.all_reduce(sm, op=torch.distributed.ReduceOp.SUM)
        ms, ms_e2e = float(mx[0]), float(mx[1])
        total_bases, U_total, launches = float(sm[2]), float(sm[3]), int(sm[4])
        h2d, d2h = int(sm[5]), int(sm[6])
    else:
        total_bases, U_total = float(my_bases), float(U_local)

    if rank == 0:
        ms_step = ms / args.steps
        value = total_bases / (ms_step * 1e-3) / 1e9
        e2e_val = total_bases / (ms_e2e / e2e_steps * 1e-3) / 1e9
        stage_ms = {k2: v / args.steps for k2, v in stage_acc.items()}
        # dominant kernel (rank 0's stage times; one launch per stage except parse/sort)
        W = stats["n_words"]
        n_windows, n_in, U0 = stats["n_windows"], stats["n_input_bytes"], stats["n_kmers"]
        stream_bytes = (stats["n_bases"] + stats["n_records"]) * 12 / 32
        # algorithmic bytes per launch.  Unit path: the packed stream is 12 B and the run masks
        # 8 B per 32 entries; a unit is 16 B; a dedupe entry (2 + WB) x 8 B; a wide record RS x 8 B.
        n_units, n_ent, n_wide = stats.get("n_units", 0), stats.get("n_unit_entries", 0), stats.get("n_wide", 0)
        mask_bytes = (stats["n_bases"] + stats["n_records"]) * 8 / 32
        WB = 1 if W <= 1 else 2 if W == 2 else 4
        ent_b, rec_b = (2 + WB) * 8.0, (2 if WB == 1 else 2 * WB) * 8.0
        if n_units:
            kernels = {
                "k_units_scatter": (stage_ms.get("scatter", 0.0), stream_bytes + mask_bytes + 16.0 * n_units),
                "k_unit_bounds": (stage_ms.get("bounds", 0.0), stream_bytes + mask_bytes),
                "k_units_dedupe": (stage_ms.get("dedupe", 0.0), 16.0 * n_units + ent_b * n_ent),
                "k_units_expand": (stage_ms.get("expand", 0.0), 2 * ent_b * n_ent + rec_b * n_wide),
                "k_aggregate_cols": (stage_ms.get("aggregate", 0.0), rec_b * n_wide + U0 * 8.0 * (1 + W)),
                "k_pack": (stage_ms.get("pack", 0.0), n_in + stream_bytes),
            }
        else:
            kernels = {
                "k_scatter": (stage_ms.get("scatter", 0.0), 8.0 * n_windows + stream_bytes),
                "k_aggregate_cols": (stage_ms.get("aggregate", 0.0), 8.0 * n_windows + U0 * 8.0 * (1 + W)),
                "k_pack": (stage_ms.get("pack", 0.0), n_in + stream_bytes),
            }
        dom = max(kernels, key=lambda n: kernels[n][0])
        dom_ms, dom_bytes = kernels[dom]
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        # whole-pipeline model of SURVEY.md 8d: B_alg = F + 8 + 8 + O bytes per base
        F = n_in / max(1, stats["n_bases"])
        O = U_total * (8 * ((G_total + 63) // 64) + 8) / total_bases
        b_alg = F + 16 + O
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic_from_profiles(dom), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": dom_ms,
                         "kernel_share_of_step": dom_ms / ms_step if ms_step else None},
            "pipeline_roofline": {"bytes_per_base": b_alg, "achieved": value * b_alg, "peak": peak,
                                  "frac": value * b_alg / peak, "unit": "GB/s",
                                  "model": "SURVEY.md 8d: F + 8 + 8 + O bytes per base"},
            "stage_ms": stage_ms,
            "kernel_gbs": {n: (b / (t * 1e-3) / 1e9 if t > 0 else None) for n, (t, b) in kernels.items()},
            "result": {"n_bases": total_bases, "n_kmers": U_total, "n_genomes": G_total,
                       "n_buckets": stats["n_buckets"], "n_splits": stats["n_splits"],
                       "n_units": n_units, "n_unit_entries": n_ent, "n_wide_records": n_wide,
                       "n_unit_buckets": stats.get("n_unit_buckets", 0)},
        }
        if world == 1 and not args.no_cpu_baseline:
            host = buf.cpu().numpy()
            n_s = min(CPU_SAMPLE_GENOMES, len(spans))
            texts = [host[o:o + n].tobytes() for o, n in spans[:n_s]]
            v, dt, threads, ref = cpu_arm(texts)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"first {n_s} genomes of the workload ({ref.n_bases} bases), "
                                              f"oracle port of multidsk+dsk2kover semantics, {dt:.1f} s"}
        print(json.dumps(line))
    db.close()
    if dist:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
