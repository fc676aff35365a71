"""Multi-GPU plumbing: samples shard over ranks (independent on a frozen tree), every rank holds the whole
flattened tree, ONE allgather of the 32-byte placement records at the end.
torch.distributed is used for the collective only (NCCL on GPUs; gloo in the CPU test-suite)."""
import numpy as np

from .capi import PLACEMENT_DTYPE

REC_WORDS = PLACEMENT_DTYPE.itemsize // 4


def shard_range(n_samples, rank, world):
    """Contiguous block of ceil(S/G) samples for rank g (the last blocks may be short or empty)."""
    per = (n_samples + world - 1) // world
    lo = min(n_samples, rank * per)
    return lo, min(n_samples, lo + per), per


def shard_batch(s_ptr, calls, rank, world):
    n = len(s_ptr) - 1
    lo, hi, per = shard_range(n, rank, world)
    sp = (s_ptr[lo:hi + 1] - s_ptr[lo]).astype(np.uint64)
    return lo, hi, per, sp, calls[int(s_ptr[lo]):int(s_ptr[hi])]


def allgather_records(local_records, n_total, world, device=None):
    """local_records: numpy PLACEMENT_DTYPE array of this rank's shard (in sample order).  Returns the
    n_total records of the whole batch on every rank.  One collective."""
    import torch
    import torch.distributed as dist
    per = (n_total + world - 1) // world
    buf = np.zeros(per, PLACEMENT_DTYPE)
    buf[: len(local_records)] = local_records
    t = torch.from_numpy(buf.view(np.int32).copy())
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * per * REC_WORDS, dtype=torch.int32, device=t.device)
    dist.all_gather_into_tensor(out, t)
    return np.frombuffer(out.cpu().numpy().tobytes(), dtype=PLACEMENT_DTYPE)[:n_total].copy()


def place_sharded(mat, s_ptr, calls, rank, world, device=None):
    """Every rank passes the same full batch; each places its contiguous shard on its own GPU, then one
    allgather returns all placements everywhere."""
    lo, hi, per, sp, sc = shard_batch(s_ptr, calls, rank, world)
    local = mat.place_batch(sp, sc)["placements"] if hi > lo else np.zeros(0, PLACEMENT_DTYPE)
    if world == 1:
        return local
    return allgather_records(local, len(s_ptr) - 1, world, device)
