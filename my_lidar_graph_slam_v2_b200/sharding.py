"""Multi-GPU sharding of a loop-detection batch.

Loop-detection queries are independent (loop_detector_branch_bound.cpp:68), so
they shard across ranks as contiguous ranges, the way the reference splits them
over its two FPGA cores (loop_detector_fpga_parallel.cpp:41-56). There is no
data-path exchange; the only collective is the 8-byte all-reduce(max) of the
packed best word

    word = key << 20 | (0xFFFFF - global query index),   key = 998*sumV + 64536*nKnown

which returns the best (score, query) pair of the whole batch on every rank:
the larger key wins, equal keys go to the lower query index. Results stay on
the rank that computed them; all_gather_results concatenates them in rank
order = query order when the caller needs the full vector (what the reference
does with its per-core result vectors).

One process per GPU; torch.distributed is only the plumbing (NCCL on the GPU
box, gloo in the CPU tests).
"""
import numpy as np

QUERY_BITS = 20
QUERY_MASK = (1 << QUERY_BITS) - 1


def shard_range(n_queries, rank, world):
    """Contiguous range [g*nq/G, (g+1)*nq/G) of rank g."""
    return (n_queries * rank) // world, (n_queries * (rank + 1)) // world


def pack_best(key, query_index):
    """Same packing as k_finalize (csrc/csm_kernels.cuh)."""
    assert 0 <= query_index <= QUERY_MASK and 0 <= key < (1 << 43)
    return (int(key) << QUERY_BITS) | (QUERY_MASK - int(query_index))


def unpack_best(word):
    word = int(word)
    if word == 0:
        return 0, -1
    return word >> QUERY_BITS, QUERY_MASK - (word & QUERY_MASK)


def local_best_word(keys, found, query_index_base):
    """Host restatement of what the device leaves in csm_best_key_device()."""
    best = 0
    for i, (k, f) in enumerate(zip(keys, found)):
        if f:
            best = max(best, pack_best(k, query_index_base + i))
    return best


def allreduce_best(word_tensor, group=None):
    """In-place all-reduce(max) of the int64 best word (NCCL or gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(word_tensor, op=dist.ReduceOp.MAX, group=group)
    return word_tensor


def all_gather_results(local_records, group=None):
    """Concatenate per-rank result records (numpy structured / 2-D array with
    equal row width) in rank order, i.e. in global query order."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local_records)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dev = None
    if dist.get_backend(group) == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        n = n.to(dev)
        counts = [c.to(dev) for c in counts]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    width = int(np.prod(local.shape[1:])) if local.ndim > 1 else 1
    flat = torch.from_numpy(local.reshape(local.shape[0], -1).view(np.uint8).copy())
    row_bytes = flat.shape[1] if local.shape[0] else width * local.dtype.itemsize
    pad = torch.zeros((max(counts), row_bytes), dtype=torch.uint8)
    pad[:flat.shape[0]] = flat
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    if dev is not None:
        pad = pad.to(dev)
        bufs = [b.to(dev) for b in bufs]
    dist.all_gather(bufs, pad, group=group)
    parts = [b[:c].cpu().numpy() for b, c in zip(bufs, counts)]
    out = np.concatenate(parts, axis=0).view(local.dtype)
    return out.reshape((sum(counts),) + local.shape[1:])
