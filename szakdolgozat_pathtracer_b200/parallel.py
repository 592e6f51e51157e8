"""Host-side logic of the multi-GPU path: the scene is replicated on every GPU, a
frame's subframes (launches) are dealt round-robin to the ranks, every rank sums its launches into a local float4
accumulator (ptb_render_cfg.accumulate_mode = 1), ONE reduce over NVLink/NVSwitch follows, the root divides by the
number of subframes and tonemaps (ptb_resolve).  Each subframe keeps its GLOBAL index, which is what seeds the RNG
(optixSphere.cu:316), so the set of samples is the same for every world size.

torch.distributed is only plumbing here (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations


def subframes_for_rank(rank: int, world: int, n_subframes: int, first: int = 0) -> list[int]:
    """Global subframe indices rank renders out of [first, first + n_subframes)."""
    if world < 1 or not (0 <= rank < world) or n_subframes < 0:
        raise ValueError("bad rank/world/n_subframes")
    return list(range(first + rank, first + n_subframes, world))


def subframe_block_for_rank(rank: int, world: int, n_subframes: int, first: int = 0) -> list[int]:
    """Contiguous split: rank r gets subframes [first + r*k, first + (r+1)*k), k = ceil(n/world) (the last ranks may get
    fewer or none).  Used when a rank renders its share as ONE batched launch (subframes_per_launch = len(result))."""
    if world < 1 or not (0 <= rank < world) or n_subframes < 0:
        raise ValueError("bad rank/world/n_subframes")
    k = -(-n_subframes // world)
    lo, hi = min(rank * k, n_subframes), min((rank + 1) * k, n_subframes)
    return list(range(first + lo, first + hi))


def resolve_scale(n_subframes: int) -> float:
    """Factor that turns the reduced sum of launch means into the frame mean."""
    if n_subframes < 1:
        raise ValueError("n_subframes must be >= 1")
    return 1.0 / float(n_subframes)


def reduce_accumulator(accum, dst: int = 0):
    """Sum the per-rank float4 accumulators onto rank dst (in place).  No-op without a process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
