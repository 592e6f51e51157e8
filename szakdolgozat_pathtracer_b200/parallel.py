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


def pixel_slice_for_rank(rank: int, world: int, n_pixels: int) -> tuple[int, int]:
    """(first, count) of the frame slice rank reduces and tonemaps in the fused peer-memory exchange (ptb_resolve_peers)."""
    if world < 1 or not (0 <= rank < world) or n_pixels < 0:
        raise ValueError("bad rank/world/n_pixels")
    k = -(-n_pixels // world)
    lo, hi = min(rank * k, n_pixels), min((rank + 1) * k, n_pixels)
    return lo, hi - lo


class PeerExchange:
    """One process per GPU: every rank exports its sum-mode accumulator (and the root its result buffers) through CUDA
    IPC; resolve() then runs ptb_resolve_peers on this rank's slice.  NCCL is used only for two stream-ordered barriers."""

    def __init__(self, ctx, rank, world, accum_ptr, out_accum_ptr, out_frame_ptr):
        import torch.distributed as dist
        self.ctx, self.rank, self.world = ctx, rank, world
        mine = dict(accum=ctx.ipc_export(accum_ptr))
        if rank == 0:
            mine.update(out_accum=ctx.ipc_export(out_accum_ptr), out_frame=ctx.ipc_export(out_frame_ptr))
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        self._opened = []
        self.accums = []
        for r in range(world):
            if r == rank:
                self.accums.append(accum_ptr)
            else:
                p = ctx.ipc_open(gathered[r]["accum"]); self._opened.append(p); self.accums.append(p)
        if rank == 0:
            self.out_accum, self.out_frame = out_accum_ptr, out_frame_ptr
        else:
            self.out_accum = ctx.ipc_open(gathered[0]["out_accum"]); self.out_frame = ctx.ipc_open(gathered[0]["out_frame"])
            self._opened += [self.out_accum, self.out_frame]

    def resolve(self, n_pixels, n_subframes, cfg, stream, barrier):
        """barrier(): a stream-ordered cross-rank barrier (e.g. a 1-element NCCL all_reduce on `stream`)."""
        barrier()  # every rank has finished rendering into its accumulator
        first, count = pixel_slice_for_rank(self.rank, self.world, n_pixels)
        self.ctx.resolve_peers(self.accums, self.out_accum, self.out_frame, first, count, resolve_scale(n_subframes), cfg, stream)
        barrier()  # every slice has landed in the root's buffers

    def close(self):
        for p in self._opened:
            try:
                self.ctx.ipc_close(p)
            except Exception:
                pass
        self._opened = []


def row_band_for_rank(rank: int, world: int, height: int) -> tuple[int, int]:
    """[row_begin, row_end) of the band rank renders under TILE partitioning (ptb_render_cfg.row_begin/row_end): the
    alternative to the sample split for single-pass frames.  Every pixel is computed whole on one GPU with its
    full-frame seed, so the tiled frame is bit-identical to the single-GPU one; there is no reduction, only a gather."""
    if world < 1 or not (0 <= rank < world) or height < 0:
        raise ValueError("bad rank/world/height")
    k = -(-height // world)
    return min(rank * k, height), min((rank + 1) * k, height)
