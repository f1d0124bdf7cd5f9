"""Worker for tests/test_distributed_cpu.py: one rank of a gloo world, oracle-backed engine."""
import os
import pickle
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch.distributed as dist  # noqa: E402

from grm_b200.distributed import DistributedBuilder, init_process_group_from_env  # noqa: E402
from tests.dist_engine import OracleEngine  # noqa: E402


def main():
    out_dir, n_genomes, k, keep = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), bool(int(sys.argv[4]))
    init_process_group_from_env("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    with open(os.path.join(out_dir, "genomes.pkl"), "rb") as f:
        genomes = pickle.load(f)
    eng = OracleEngine(k, 1, keep)
    db = DistributedBuilder(k=k, keep_singletons=keep, n_genomes=n_genomes, rank=rank, world=world, engine=eng, builder=eng)
    db.reset()
    for i, g in enumerate(db.local_rows):
        db.add_genome_bytes(i, genomes[g])
    db.build()
    res = db.gather()
    # the host-side counts exchange of the peer path (distributed.HostCounts), several rounds in both buffer halves
    from grm_b200.distributed import HostCounts
    if HostCounts.available(world):
        hc = HostCounts(dist, world, rank)
        for rnd in range(1, 6):
            M = hc.exchange([1000 * rnd + 10 * rank + d for d in range(world)])
            assert M == [[1000 * rnd + 10 * s + d for d in range(world)] for s in range(world)], M
    if rank == 0:
        np.savez(os.path.join(out_dir, "result.npz"), kmers=res[0], matrix=res[1])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
