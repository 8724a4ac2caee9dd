"""Multi-GPU plumbing: clips are independent, so a batch is split contiguously over ranks (one process per GPU) and
every rank decodes its own shard with no data-path collective.  The only communication is an
optional final gather of the int64 paths / float32 scores to rank 0 (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_bounds(n_clips, rank, world_size):
    """Contiguous shard [lo, hi) of rank `rank`: sizes differ by at most one clip, earlier ranks get the extra ones."""
    assert 0 <= rank < world_size
    base, rem = divmod(int(n_clips), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def decode_sharded(decode_fn, log_emis, lengths=None, rank=0, world_size=1, gather=False, group=None):
    """Decode this rank's shard of `log_emis [B, T, S]` with `decode_fn(emis_shard, lengths_shard) -> (paths, scores)`.

    gather=False: returns (lo, hi, paths_shard, scores_shard).
    gather=True : rank 0 returns the full (paths [B, T], scores [B]) in clip order, other ranks return (None, None);
                  uses torch.distributed.gather_object-free tensor gathers on `group` (backend decides the device).
    """
    B = log_emis.shape[0]
    lo, hi = shard_bounds(B, rank, world_size)
    sub_len = None if lengths is None else lengths[lo:hi]
    paths, scores = decode_fn(log_emis[lo:hi], sub_len)
    if not gather:
        return lo, hi, paths, scores

    import torch
    import torch.distributed as dist
    T = log_emis.shape[1]
    dev = paths.device if torch.is_tensor(paths) else torch.device('cpu')
    p = torch.as_tensor(paths).to(dev).contiguous()
    s = torch.as_tensor(scores).to(dev).contiguous()
    # pad every shard to the largest shard so that one fixed-shape gather works on every backend
    nmax = shard_bounds(B, 0, world_size)[1]
    pp = torch.full((nmax, T), -1, dtype=torch.int64, device=dev)
    ss = torch.full((nmax,), float('-inf'), dtype=torch.float32, device=dev)
    pp[:hi - lo] = p
    ss[:hi - lo] = s
    if rank == 0:
        plist = [torch.empty_like(pp) for _ in range(world_size)]
        slist = [torch.empty_like(ss) for _ in range(world_size)]
    else:
        plist = slist = None
    dist.gather(pp, plist, dst=0, group=group)
    dist.gather(ss, slist, dst=0, group=group)
    if rank != 0:
        return None, None
    out_p = torch.empty((B, T), dtype=torch.int64, device=dev)
    out_s = torch.empty((B,), dtype=torch.float32, device=dev)
    for r in range(world_size):
        a, b = shard_bounds(B, r, world_size)
        out_p[a:b] = plist[r][:b - a]
        out_s[a:b] = slist[r][:b - a]
    if not torch.is_tensor(paths):
        return out_p.cpu().numpy(), out_s.cpu().numpy()
    return out_p, out_s
