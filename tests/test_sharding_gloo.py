"""CPU, world_size 2 over gloo: the N > 1 path of bench.py (row blocks per rank, no data-path
collective, MAX-reduce of timings).  The per-rank compute here is the oracle -- this test is
about the host-side sharding logic, not the kernels."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import torch.distributed as dist
    from oracle import oracle as O
    from sharding import aggregate_throughput, max_over_ranks, row_block

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = np.cumsum(np.random.default_rng(1).standard_normal((11, 30)), axis=1)
    y = np.cumsum(np.random.default_rng(2).standard_normal((7, 30)), axis=1)
    lo, hi = row_block(len(x), world, rank)
    slab = O.pairwise("dtw", x[lo:hi], y, r=0.2)
    np.save(os.path.join(tmp, f"slab{rank}.npy"), slab)
    np.save(os.path.join(tmp, f"range{rank}.npy"), np.array([lo, hi]))
    t = max_over_ranks([10.0 + rank, 5.0 - rank])
    assert t == [10.0 + world - 1, 5.0], t
    assert aggregate_throughput(100.0, 2, t[0]) == 100.0 * 2 / (t[0] * 1e-3)
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharding_world_size_2(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    x = np.cumsum(np.random.default_rng(1).standard_normal((11, 30)), axis=1)
    y = np.cumsum(np.random.default_rng(2).standard_normal((7, 30)), axis=1)
    full = np.empty((11, 7))
    covered = np.zeros(11, dtype=int)
    for r in range(world):
        lo, hi = np.load(tmp_path / f"range{r}.npy")
        full[lo:hi] = np.load(tmp_path / f"slab{r}.npy")
        covered[lo:hi] += 1
    assert (covered == 1).all()
    assert np.array_equal(full, oracle.pairwise("dtw", x, y, r=0.2))


def test_row_block_partition():
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from sharding import row_block
    for n in (1, 7, 8, 10000, 10001):
        for nb in (1, 2, 3, 4, 8):
            if nb > n:
                continue
            blocks = [row_block(n, nb, b) for b in range(nb)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(nb - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
