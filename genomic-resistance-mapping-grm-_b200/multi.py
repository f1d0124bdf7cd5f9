"""N GPUs behind the reference-facing wrappers (``from_contigs`` / ``from_reads`` / the CLI / the Surveyor runner).

``GRM_GPUS=N`` (or ``gpus=N``) makes a dataset build run as N processes, one per GPU, the way the GUI launches Ray with
``mpiexec -n 4`` (src/app.py:1310): the wrapper writes the job (files per genome row, k, filters) to a JSON file and starts
``python -m torch.distributed.run --nproc-per-node N -m grm_b200.multi job.json``.  Every rank builds the rows of its
64-aligned block (distributed.py), the ranks exchange partial columns by hash range, and every rank leaves ITS slice of
the columns (k-mer strings + all matrix word rows) as ``slice_<rank>.*.npy`` next to the job file -- or, for a Surveyor
job, writes its rows of ``KmerMatrix.tsv`` at their fixed offset.  The slices in rank order are the one-GPU column order
(ascending hash), so the parent just concatenates them while it writes the HDF5 file.  Nothing is pickled through
``gather_object`` and no rank ever holds the whole matrix.

A process that already runs under torchrun (WORLD_SIZE set) joins as a rank instead of spawning.
"""
from __future__ import annotations

import json
import os
import socket
import subprocess
import sys
import tempfile

import numpy as np


def requested_gpus(gpus=None) -> int:
    if gpus is None:
        gpus = os.environ.get("GRM_GPUS", "1")
    try:
        return max(1, int(gpus))
    except (TypeError, ValueError):
        return 1


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_rank(job: dict, job_dir: str) -> None:
    """One rank of the job (under torchrun)."""
    import torch
    from .distributed import DistributedBuilder, init_process_group_from_env
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    backend = job.get("backend")
    if torch.cuda.device_count() < int(os.environ.get("LOCAL_WORLD_SIZE", world)):
        backend = "gloo"                     # more ranks than GPUs (tests): CUDA engine, host transport
    device = local_rank % max(1, torch.cuda.device_count())
    torch.cuda.set_device(device)
    dist = init_process_group_from_env(backend)
    files = job["files"]
    db = DistributedBuilder(k=job["k"], min_abundance=job["min_abundance"], keep_singletons=job["keep_singletons"],
                            n_genomes=len(files), rank=rank, world=world, input_kind=job["input_kind"], device=device)
    try:
        db.reset()
        for i, g in enumerate(db.local_rows):
            if files[g]:
                db.add_genome_files(i, files[g])
        db.build()
        if job.get("tsv"):
            n = db.write_tsv(job["tsv"], job["names"])
            if rank == 0:
                print("[Surveyor] %d samples, %d k-mers -> %s (%d GPUs)" % (len(files), n, job["tsv"], world))
        else:
            counts = db.slice_counts()
            seqs = db.builder.kmer_strings() if db.n_kmers else np.zeros(0, dtype="S%d" % job["k"])
            _, mat = db.result_host()
            np.save(os.path.join(job_dir, "slice_%d.kmers.npy" % rank), seqs)
            np.save(os.path.join(job_dir, "slice_%d.matrix.npy" % rank), np.asarray(mat))
            if rank == 0:
                st = dict(db.local_stats)
                with open(os.path.join(job_dir, "result.json"), "w") as f:
                    json.dump({"world": world, "counts": counts, "n_bases_rank0": st.get("n_bases", 0)}, f)
        dist.barrier()
    finally:
        db.close()
        dist.destroy_process_group()


def launch(job: dict, gpus: int, temp_dir=None) -> str:
    """Run the job on `gpus` ranks; returns the job directory (slices / result.json inside).  Raises on failure."""
    job_dir = tempfile.mkdtemp(prefix="grmkm_job_", dir=temp_dir if temp_dir and os.path.isdir(str(temp_dir)) else None)
    job_path = os.path.join(job_dir, "job.json")
    with open(job_path, "w") as f:
        json.dump(job, f)
    if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1:
        _run_rank(job, job_dir)                    # already a rank of somebody's torchrun: join
        return job_dir
    env = dict(os.environ)
    env.pop("GRM_GPUS", None)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(gpus),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), "-m", "grm_b200.multi", job_path]
    p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        raise RuntimeError("the %d-GPU build failed (exit status %d):\n%s" % (gpus, p.returncode, p.stdout[-4000:]))
    for line in p.stdout.splitlines():
        if line.startswith("[Surveyor]") or line.startswith("grm_b200:"):
            print(line, flush=True)
    return job_dir


def build_matrix(files_per_genome, k, min_abundance, keep_singletons, input_kind, gpus, temp_dir=None):
    """-> (kmer_strings S<k>[U], matrix uint64[W][U], stats) assembled from the ranks' slices (rank order = hash order)."""
    job = {"files": [list(f) for f in files_per_genome], "k": int(k), "min_abundance": int(min_abundance),
           "keep_singletons": bool(keep_singletons), "input_kind": int(input_kind)}
    job_dir = launch(job, gpus, temp_dir)
    try:
        with open(os.path.join(job_dir, "result.json")) as f:
            res = json.load(f)
        seqs = [np.load(os.path.join(job_dir, "slice_%d.kmers.npy" % r)) for r in range(res["world"])]
        mats = [np.load(os.path.join(job_dir, "slice_%d.matrix.npy" % r), mmap_mode="r") for r in range(res["world"])]
        kmer_strings = np.concatenate(seqs) if seqs else np.zeros(0, dtype="S%d" % k)
        W = (len(files_per_genome) + 63) // 64
        matrix = np.concatenate([m.reshape(W, -1) for m in mats], axis=1) if mats else np.zeros((W, 0), dtype=np.uint64)
        return kmer_strings, np.ascontiguousarray(matrix), {"n_kmers": int(matrix.shape[1]), "n_gpus": res["world"]}
    finally:
        cleanup(job_dir)


def cleanup(job_dir: str) -> None:
    for name in os.listdir(job_dir):
        try:
            os.remove(os.path.join(job_dir, name))
        except OSError:
            pass
    try:
        os.rmdir(job_dir)
    except OSError:
        pass


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        print("usage (under torchrun): python -m grm_b200.multi job.json", file=sys.stderr)
        return 2
    with open(argv[0]) as f:
        job = json.load(f)
    _run_rank(job, os.path.dirname(os.path.abspath(argv[0])))
    return 0


if __name__ == "__main__":
    sys.exit(main())
