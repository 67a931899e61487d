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
