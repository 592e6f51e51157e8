"""N > 1 host logic on CPU: two gloo ranks split the subframes of a frame exactly as bench.py does on GPUs
(szakdolgozat_pathtracer_b200/parallel.py), render them with the oracle in sum mode, reduce, resolve; the result must
equal the single-process sum over the same global subframe indices."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent

W, H, SUBFRAMES = 48, 32, 4
KW = dict(spp_per_launch=2, max_depth=4, accumulate_sum=1)


def _setup():
    for p in (ROOT, ROOT / "tests", ROOT / "tools"):
        if str(p) not in sys.path:
            sys.path.insert(0, str(p))
    import make_assets
    import orchelp as oh
    import szakdolgozat_pathtracer_b200 as ptb
    from scenes import load_config
    sc = load_config(ptb, make_assets, "c1", small=True)
    return ptb, oh, oh.OracleScene.from_ptb(sc, guard=False)


def _render_subframes(ptb, oh, osc, subs):
    accum = np.zeros((H, W, 4), np.float32)
    for sf in subs:
        p = ptb.make_params(W, H, subframe_index=sf, dof=True)
        accum, _, _, _, rc = oh.render("oracle", osc, oh.params_from_ptb(p), oh.default_config("oracle", threads=1, **KW), accum=accum, want_hits=False)
        assert rc == 0
    return accum


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ptb, oh, osc = _setup()
    from szakdolgozat_pathtracer_b200 import parallel
    subs = parallel.subframes_for_rank(rank, world, SUBFRAMES)
    t = torch.from_numpy(_render_subframes(ptb, oh, osc, subs))
    parallel.reduce_accumulator(t, dst=0)
    if rank == 0:
        np.save(out_path, t.numpy()[..., :3] * np.float32(parallel.resolve_scale(SUBFRAMES)))
    dist.barrier()
    dist.destroy_process_group()


def test_subframe_schedule():
    from szakdolgozat_pathtracer_b200 import parallel
    for world in (1, 2, 4, 8):
        got = sorted(s for r in range(world) for s in parallel.subframes_for_rank(r, world, 64, first=8))
        assert got == list(range(8, 72))
    assert parallel.subframes_for_rank(1, 2, 5) == [1, 3]
    assert parallel.subframes_for_rank(3, 4, 2) == []  # ragged: more ranks than subframes
    assert parallel.resolve_scale(8) == 0.125
    for world in (1, 2, 3, 8):
        blocks = [parallel.subframe_block_for_rank(r, world, 17, first=4) for r in range(world)]
        assert sum(blocks, []) == list(range(4, 21))
        bands = [parallel.row_band_for_rank(r, world, 1081) for r in range(world)]
        assert bands[0][0] == 0 and bands[-1][1] == 1081 and all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
        slices = [parallel.pixel_slice_for_rank(r, world, 1920 * 1080 + 5) for r in range(world)]
        assert slices[0][0] == 0 and sum(c for _, c in slices) == 1920 * 1080 + 5
        assert all(slices[i][0] + slices[i][1] == slices[i + 1][0] for i in range(world - 1))
    import pytest
    with pytest.raises(ValueError):
        parallel.subframes_for_rank(2, 2, 4)


def test_two_rank_sample_split_equals_single_process(tmp_path):
    out = tmp_path / "reduced.npy"
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    ptb, oh, osc = _setup()
    ref = _render_subframes(ptb, oh, osc, range(SUBFRAMES))[..., :3] * np.float32(1.0 / SUBFRAMES)
    got = np.load(out)
    # summation order differs (0+2)+(1+3) vs ((0+1)+2)+3: equal to rounding, not bitwise
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-6)
    assert got.mean() > 0.01
