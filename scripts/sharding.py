"""Row sharding helpers shared by bench.py and the multi-process tests (bench tooling, not part of the product package).

The distance matrix shards by contiguous row blocks of x (the reference's thread partitioner,
utils/_parallel.py:7-23); y is replicated; every rank writes only its own rows, so the data path
needs no collective.  The only cross-rank exchange is the MAX-reduce of the per-rank timings.
"""


def row_block(n, n_blocks, b):
    """[lo, hi) of block b when n rows are cut into n_blocks contiguous blocks
    (the first n % n_blocks blocks are one row longer)."""
    bs, ov = divmod(n, n_blocks)
    lo = b * bs + min(b, ov)
    return lo, lo + bs + (1 if b < ov else 0)


def max_over_ranks(values, device=None):
    """Element-wise MAX of a list of floats over all ranks (identity when not distributed)."""
    import torch
    import torch.distributed as dist

    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def aggregate_throughput(units_all_ranks, steps, ms_max):
    """Whole-job throughput: units processed by ALL ranks over the slowest rank's time."""
    return units_all_ranks * steps / (ms_max * 1e-3)
