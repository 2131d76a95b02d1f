"""N>1 host logic on CPU (gloo, world_size 2): the (tile, CFG-branch) unit partition of
stableanimator_b200/scheduler.py plus one all-reduce(sum) of the weighted accumulator per step must
reproduce the single-process tile blend of inference_pipeline_animation.py:656-689."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stableanimator_b200 import scheduler as sch


def fake_eps(tile_idx, branch, step, frames_idx, h, w):
    g = torch.Generator().manual_seed(1000 * step + 10 * tile_idx + branch)
    return torch.randn(len(frames_idx), 4, h, w, generator=g)


def accumulate(units, tiles, weights, F, h, w, step):
    noise = torch.zeros(2, F, 4, h, w)
    for t, b in units:
        idx = tiles[t]
        noise[b, idx] += fake_eps(t, b, step, idx, h, w) * weights[:, None, None, None]
    return noise


def _worker(rank, world, store_path, F, tile, overlap, q):
    # file rendezvous: no TCP port to collide with another test run on the same host
    dist.init_process_group("gloo", init_method=f"file://{store_path}", rank=rank, world_size=world)
    tiles = sch.tile_indices(F, tile, overlap)
    weights = torch.tensor(sch.tile_weights(tile))
    mine = sch.partition_units(len(tiles), world)[rank]
    out = []
    for step in range(2):
        noise = accumulate(mine, tiles, weights, F, 4, 4, step)
        dist.all_reduce(noise)
        out.append(noise)
    if rank == 0:
        q.put(torch.stack(out))
    dist.barrier()
    dist.destroy_process_group()


def _run_world2(tmp_path, attempt, F, tile, overlap):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    store = str(tmp_path / f"rendezvous{attempt}")
    procs = [ctx.Process(target=_worker, args=(r, 2, store, F, tile, overlap, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        got = q.get(timeout=300)
    except Exception:  # a worker died before reporting (seen once under heavy host load)
        got = None
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.kill()
    # rank 0 reported after the all-reduces: the result stands (and is checked against the single-process blend below) even
    # if a rank then trips over gloo's teardown on a loaded host
    return got


@pytest.mark.parametrize("F,tile,overlap", [(16, 16, 4), (64, 16, 4), (128, 16, 4)])
def test_partitioned_blend_equals_single_process(F, tile, overlap, tmp_path):
    got = None
    for attempt in range(3):  # retries: process spawn + gloo rendezvous on a loaded CI host
        got = _run_world2(tmp_path, attempt, F, tile, overlap)
        if got is not None:
            break
    assert got is not None, "world-size-2 gloo run failed three times"
    tiles = sch.tile_indices(F, tile, overlap)
    weights = torch.tensor(sch.tile_weights(tile))
    units = [(t, b) for t in range(len(tiles)) for b in (0, 1)]
    for step in range(2):
        ref = accumulate(units, tiles, weights, F, 4, 4, step)
        assert torch.allclose(got[step], ref, atol=1e-6)


def test_partition_properties():
    for n_tiles in (1, 2, 5, 11):
        for world in (1, 2, 4, 8):
            parts = sch.partition_units(n_tiles, world)
            flat = [u for p in parts for u in p]
            assert sorted(flat) == [(t, b) for t in range(n_tiles) for b in (0, 1)]
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1
    # granularity ceilings quoted in SURVEY 8e: 10 units on 4 ranks -> 3 rounds, 22 on 8 -> 3 rounds
    assert max(len(p) for p in sch.partition_units(5, 4)) == 3
    assert max(len(p) for p in sch.partition_units(11, 8)) == 3
    assert sch.partition_units(1, 2) == [[(0, 0)], [(0, 1)]]  # CFG split (BASELINE configs[2])


def test_tile_indices_rejects_short_videos():
    with pytest.raises(ValueError):
        sch.tile_indices(8, 16, 4)
