"""Multi-GPU plumbing: one process per GPU, ciphertext batches partitioned contiguously over ranks, keys distributed
once.  There is no data-path collective: the only exchanges are the start-up key broadcast and an
optional gather of the outputs.  Backend-agnostic (NCCL on the GPU box, gloo in the CPU test tier)."""
import numpy as np


def shard_range(count, rank, world):
    """Contiguous split of range(count): the first count % world ranks get one extra item."""
    base, rem = divmod(count, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_host_arrays(dist, arrays, root=0):
    """Start-up distribution of key material held as host numpy arrays (reference layout) from `root`.
    Non-root ranks pass arrays of the right shape/dtype (contents ignored)."""
    import torch
    out = []
    for a in arrays:
        t = torch.from_numpy(np.ascontiguousarray(a).view(np.int64 if a.dtype == np.uint64 else a.dtype).copy())
        dist.broadcast(t, src=root)
        out.append(t.numpy().view(a.dtype))
    return out


def sharded_map(dist, fn, batch, out_width=None, gather=True):
    """Apply `fn` (rows -> rows, e.g. a batched bootstrap) to this rank's contiguous shard of `batch` [count, width];
    with gather=True every rank returns the full result in the original order (one all_gather of padded shards)."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    count = batch.shape[0]
    lo, hi = shard_range(count, rank, world)
    mine = fn(batch[lo:hi]) if hi > lo else np.zeros((0, out_width or batch.shape[1]), dtype=batch.dtype)
    if not gather:
        return mine
    width = mine.shape[1] if mine.size else (out_width or batch.shape[1])
    pad = (count + world - 1) // world
    buf = np.zeros((pad, width), dtype=np.int64)
    buf[: hi - lo] = mine.view(np.int64) if mine.dtype == np.uint64 else mine
    parts = [torch.zeros((pad, width), dtype=torch.int64) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(buf))
    rows = []
    for r in range(world):
        rlo, rhi = shard_range(count, r, world)
        rows.append(parts[r].numpy()[: rhi - rlo])
    full = np.concatenate(rows, axis=0)
    return full.view(np.uint64) if batch.dtype == np.uint64 else full
