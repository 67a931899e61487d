"""Multi-GPU plumbing for the verify path (one process per GPU, torch.distributed).

Every (pk, msg, sig) triple is independent, so a batch is split into contiguous index ranges, one per rank, and the
data path needs no collective.  The single exchange step at the end is an all-gather of (a) each rank's packed
ok-bitmap shard and (b) each rank's 576-byte GT partial product, followed by a local fold of the partials in rank
order (Fp12 multiplication is not a reduction operator NCCL knows, hence all-gather + fold rather than all-reduce;
the product is commutative, so every rank obtains identical bytes)."""
import torch
import torch.distributed as dist

def shard_range(n, world, rank):
    """contiguous shard [lo, hi) of rank; shard sizes are multiples of 64 (except the last) so bitmap words never straddle ranks"""
    per = -(-n // world); per = -(-per // 64) * 64
    lo = min(n, rank * per); hi = min(n, lo + per)
    return lo, hi

def shard_words(n, world):
    lo, hi = shard_range(n, world, 0)
    return (hi - lo + 63) // 64

def exchange(bitmap_shard, gt_partial, fold, group=None):
    """bitmap_shard: int64[words] (padded to shard_words), gt_partial: uint8[576]; fold(parts uint8[world*576]) -> uint8[576].
    Returns (full bitmap int64[world*words], folded GT uint8[576]).  Works on any backend (nccl on device tensors, gloo on CPU)."""
    world = dist.get_world_size(group)
    bm = torch.empty(world * bitmap_shard.numel(), dtype=bitmap_shard.dtype, device=bitmap_shard.device)
    gt = torch.empty(world * 576, dtype=torch.uint8, device=gt_partial.device)
    dist.all_gather_into_tensor(bm, bitmap_shard, group=group)
    dist.all_gather_into_tensor(gt, gt_partial, group=group)
    return bm, fold(gt)


def witness_shard(nwit, world, rank):
    """R1CS check: the matrices are replicated, the assignments are split into contiguous equal shards
    (the last ranks may hold one fewer); returns [lo, hi)"""
    base, extra = divmod(nwit, world)
    lo = rank * base + min(rank, extra); hi = lo + base + (1 if rank < extra else 0)
    return lo, hi

def gather_flags(all_sat_shard, nwit, group=None):
    """all_sat_shard: uint8[hi - lo] of this rank; returns uint8[nwit] on every rank (one all-gather of padded shards --
    the per-constraint bit vectors stay with the rank that owns the assignment)"""
    world = dist.get_world_size(group); per = -(-nwit // world)
    pad = torch.zeros(per, dtype=torch.uint8, device=all_sat_shard.device); pad[:all_sat_shard.numel()] = all_sat_shard
    out = torch.empty(world * per, dtype=torch.uint8, device=all_sat_shard.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = []
    for r in range(world):
        lo, hi = witness_shard(nwit, world, r); parts.append(out[r * per: r * per + (hi - lo)])
    return torch.cat(parts)
