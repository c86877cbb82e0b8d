"""Multi-GPU plumbing of the hot path: one process per GPU, the samples of every pixel split across
ranks, one reduce of the fp32 radiance buffer to rank 0 per frame.  No other collective exists on the path.

The sample streams are keyed by the GLOBAL sample index, so the frame is the same for any number of ranks."""
from __future__ import annotations


def sample_range(rank: int, world: int, spp_total: int) -> tuple[int, int]:
    """(sample_begin, sample_count) of `rank`: contiguous blocks, the remainder spread over the first ranks."""
    if world <= 0 or not (0 <= rank < world) or spp_total < 0:
        raise ValueError("bad rank / world / spp")
    base, rem = divmod(spp_total, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def reduce_frame(fb, dist=None, dst: int = 0):
    """Sums the per-rank partial frames (each already divided by the frame's total spp) onto `dst`.
    `fb` is a torch tensor on the rank's device (NCCL) or on the CPU (gloo, tests)."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(fb, dst=dst, op=dist.ReduceOp.SUM)
    return fb
