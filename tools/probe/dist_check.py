"""torchrun check on real GPUs: the N-rank build (peer exchange or NCCL all-to-all, GRM_EXCHANGE) gathers to the same
bytes as a one-GPU build.  usage: torchrun --nproc-per-node N tools/probe/dist_check.py [n_genomes] [scale]"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from grm_b200 import synth
from grm_b200.builder import KmerMatrixBuilder
from grm_b200.distributed import DistributedBuilder, init_process_group_from_env

G = int(sys.argv[1]) if len(sys.argv) > 1 else 150
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.02
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
init_process_group_from_env()
cfg = synth.SynthConfig(seed=synth.MASTER_SEED + 7).scaled(scale)
for keep in (True, False):
    db = DistributedBuilder(k=31, keep_singletons=keep, n_genomes=G, rank=rank, world=world, device=lr)
    for rep in range(2):                       # the second build reuses the symmetric buffer and the partition
        db.reset()
        texts = [synth.genome_fasta(cfg, g) for g in db.local_rows]
        db.add_genomes(list(range(len(texts))), texts)
        db.build(reuse_partition=rep > 0)
    res = db.gather()
    if rank == 0:
        with KmerMatrixBuilder(k=31, keep_singletons=keep, device=lr) as b:
            b.add_genomes(list(range(G)), [synth.genome_fasta(cfg, g) for g in range(G)])
            b.build()
            km, mat = b.kmers(), b.matrix()
        ok = np.array_equal(res[0], km) and np.array_equal(res[1], mat)
        print(f"world {world} exchange {'peer' if db._peer else 'nccl'} keep {keep}: {len(km)} columns, identical to one GPU: {ok}", flush=True)
        assert ok
    db.close()
dist.barrier()
dist.destroy_process_group()
