"""N>1 host logic on CPU: world_size-2 (and 3) gloo groups, oracle-backed engine, vs single-process oracle."""
import os
import pickle
import socket
import subprocess
import sys

import numpy as np
import pytest

from grm_b200.distributed import row_partition, words_per_rank
from oracle import oracle
from tests import inputs

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_row_partition_is_64_aligned_and_complete():
    for G, P in [(1000, 8), (250, 2), (500, 4), (100, 1), (65, 2), (64, 2), (10, 4), (0, 2), (130, 3)]:
        parts = row_partition(G, P)
        assert [g for r in parts for g in r] == list(range(G))
        assert all(r.start % 64 == 0 or len(r) == 0 for r in parts)
        assert sum(words_per_rank(G, P)) == (G + 63) // 64
        for r, w in zip(parts, words_per_rank(G, P)):
            assert (len(r) + 63) // 64 == w
    assert [len(r) for r in row_partition(1000, 8)] == [128] * 7 + [104]


@pytest.mark.parametrize("world,G,keep", [(2, 130, False), (2, 70, True), (3, 200, False)])
def test_gloo_world_matches_oracle(tmp_path, world, G, keep):
    rng = np.random.default_rng(world * 100 + G)
    shared = [inputs.rand_seq(rng, 300)]
    genomes = [inputs.fasta(rng, n_records=2, max_len=120, shared=shared) for _ in range(G)]
    with open(tmp_path / "genomes.pkl", "wb") as f:
        pickle.dump(genomes, f)
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_dist_worker.py"), str(tmp_path), str(G), "11",
                                       str(int(keep))], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out.decode()[-2000:]
    got = np.load(tmp_path / "result.npz")
    ref = oracle.build([[(g, 0)] for g in genomes], 11, 1, keep)
    assert np.array_equal(got["kmers"], ref.kmers)
    assert np.array_equal(got["matrix"], ref.matrix)
