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
    """One process per GPU (torchrun): the fused peer-memory exchange of a sample-split (or tile-split) frame, ordered by
    epoch flags in peer-mapped device memory instead of NCCL barriers (include/ptb.h: ptb_peer_signal,
    ptb_resolve_peers_sync, ptb_peer_wait).

    Every rank owns TWO sum-mode accumulators (steps alternate between them) and one flag block; the root also owns the
    result buffers.  All of it is exported through CUDA IPC once, at set-up (torch.distributed is used for that
    all-gather only).  One step on rank r:
        acc = begin_step()            the accumulator of this step (zero it, render into it)
        resolve(...)                  signal "rendered" to every rank -> ONE kernel: wait for all ranks' signals, reduce
                                      this rank's slice of the frame out of all accumulators (NVLink peer loads), tonemap,
                                      store into the root's buffers, last block signals "slice landed" to the root
                                      -> root only: device-side wait for all slices.
    Why two accumulators are enough: rank r re-uses accumulator A two steps later, after its resolve of the step in between,
    which waited for every peer's "rendered" signal of that step -- and a peer sends that only after its own resolve of
    the earlier step (the one that read A) has finished, by stream order."""

    def __init__(self, ctx, rank, world, n_pixels):
        import torch.distributed as dist
        self.ctx, self.rank, self.world, self.n_pixels = ctx, rank, world, n_pixels
        self.epoch = 0
        self.accum = [ctx.alloc(n_pixels * 16), ctx.alloc(n_pixels * 16)]
        self.flags = ctx.peer_flags_create()
        mine = dict(accum0=ctx.ipc_export(self.accum[0]), accum1=ctx.ipc_export(self.accum[1]), flags=ctx.ipc_export(self.flags))
        self._owned = list(self.accum) + [self.flags]
        if rank == 0:
            self.out_accum, self.out_frame = ctx.alloc(n_pixels * 16), ctx.alloc(n_pixels * 4)
            self._owned += [self.out_accum, self.out_frame]
            mine.update(out_accum=ctx.ipc_export(self.out_accum), out_frame=ctx.ipc_export(self.out_frame))
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        self._opened = []

        def opened(handle):
            ptr = ctx.ipc_open(handle)
            self._opened.append(ptr)
            return ptr
        self.accums = [[self.accum[b] if r == rank else opened(gathered[r][f"accum{b}"]) for r in range(world)] for b in (0, 1)]
        self.flag_blocks = [self.flags if r == rank else opened(gathered[r]["flags"]) for r in range(world)]
        if rank != 0:
            self.out_accum, self.out_frame = opened(gathered[0]["out_accum"]), opened(gathered[0]["out_frame"])

    def begin_step(self) -> int:
        """Device pointer of the accumulator this step renders into."""
        self.epoch += 1
        return self.accum[self.epoch & 1]

    def resolve(self, n_subframes, cfg, stream):
        """After this step's rendering has been enqueued on `stream`."""
        e = self.epoch
        self.ctx.peer_signal(self.flag_blocks, self.rank, 0, e, stream)
        first, count = pixel_slice_for_rank(self.rank, self.world, self.n_pixels)
        self.ctx.resolve_peers_sync(self.accums[e & 1], self.rank, self.flags, self.flag_blocks[0], e, self.out_accum, self.out_frame,
                                    first, count, resolve_scale(n_subframes), cfg, stream)
        if self.rank == 0:
            self.ctx.peer_wait(self.flags, 1, self.world, e, stream)

    def timed_out(self, stream=0) -> bool:
        return self.ctx.peer_flags_error(self.flags, stream)

    def close(self):
        for p in self._opened:
            try:
                self.ctx.ipc_close(p)
            except Exception:
                pass
        self._opened = []
        for p in self._owned:
            try:
                self.ctx.free(p)
            except Exception:
                pass
        self._owned = []


def row_band_for_rank(rank: int, world: int, height: int) -> tuple[int, int]:
    """[row_begin, row_end) of the band rank renders under TILE partitioning (ptb_render_cfg.row_begin/row_end): the
    alternative to the sample split for single-pass frames.  Every pixel is computed whole on one GPU with its
    full-frame seed, so the tiled frame is bit-identical to the single-GPU one; there is no reduction, only a gather."""
    if world < 1 or not (0 <= rank < world) or height < 0:
        raise ValueError("bad rank/world/height")
    k = -(-height // world)
    return min(rank * k, height), min((rank + 1) * k, height)
