"""N > 1 host logic on CPU: tile sharding plan + gloo all-reduce of the boundary table (world_size 2 and 3).

The per-rank partial accumulators are produced by the C oracle here (there is no GPU in the CPU suite);
what is under test is the product's sharding / slot / merge logic in proj_roadsurf_b200.distributed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cport
from proj_roadsurf_b200 import synth
from proj_roadsurf_b200.distributed import global_rows, merge_boundary, plan_shards, tile_ranges


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    g = synth.Grid(6, 8)
    rr = synth.ribbon_roads(g, 60, seed=12)
    tiles = synth.host_tiles(g, 3)
    return g, rr, tiles


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, rr, tiles = _case()
    sh = plan_shards(rr.roads, rr.pairs, g.n_tiles, world, only_rank=rank)[rank]
    gt = g.transforms(np.arange(sh.tile_lo, sh.tile_hi))
    h, z = cport.zonal_accumulate(sh.roads.xy, sh.roads.ring_off, sh.roads.road_ring_off, sh.pairs.road_pair_off,
                                  sh.pairs.pair_tile, tiles[sh.tile_lo:sh.tile_hi], gt)
    hist = torch.zeros((sh.n_rows, 3, 256), dtype=torch.int32)
    nz = torch.zeros((sh.n_rows,), dtype=torch.int32)
    hist[torch.from_numpy(sh.slot.astype(np.int64))] = torch.from_numpy(h.astype(np.int32))
    nz[torch.from_numpy(sh.slot.astype(np.int64))] = torch.from_numpy(z.astype(np.int32))
    merge_boundary(hist, nz, sh.n_own)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), hist=hist.numpy(), nz=nz.numpy(), rows=global_rows(sh), n_own=sh.n_own)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_merge_equals_single_shot(tmp_path, world):
    g, rr, tiles = _case()
    full_h, full_z = cport.zonal_accumulate(rr.roads.xy, rr.roads.ring_off, rr.roads.road_ring_off,
                                            rr.pairs.road_pair_off, rr.pairs.pair_tile, tiles, g.transforms())
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    seen = np.zeros(60, int)
    n_boundary = None
    for r in range(world):
        d = np.load(tmp_path / f"r{r}.npz")
        rows, n_own = d["rows"], int(d["n_own"])
        assert np.array_equal(d["hist"].view(np.uint32).astype(np.uint64), full_h[rows])      # own AND boundary rows final
        assert np.array_equal(d["nz"].view(np.uint32).astype(np.uint64), full_z[rows])
        seen[rows[:n_own]] += 1
        n_boundary = len(rows) - n_own
        bnd = rows[n_own:]
    assert n_boundary > 0, "the case must have roads crossing shard borders"
    seen[bnd] += 1
    has_pairs = np.diff(rr.pairs.road_pair_off) > 0
    assert np.array_equal(seen > 0, has_pairs) and seen.max() == 1           # every road final exactly once


def test_plan_invariants():
    g, rr, _ = _case()
    world = 4
    cuts = tile_ranges(g.n_tiles, world)
    assert cuts[0] == 0 and cuts[-1] == g.n_tiles
    shards = plan_shards(None, rr.pairs, g.n_tiles, world)
    total_pairs = sum(s.pairs.n_pairs for s in shards)
    assert total_pairs == rr.pairs.n_pairs
    for s in shards:
        assert s.pairs.pair_tile.min() >= 0 and s.pairs.pair_tile.max() < s.tile_hi - s.tile_lo
        assert len(np.unique(s.slot)) == len(s.slot) and s.slot.max() < s.n_rows
        assert np.array_equal(s.boundary_global, shards[0].boundary_global)


def test_balanced_cuts_level_the_pair_counts():
    from proj_roadsurf_b200.distributed import balanced_tile_ranges
    g = synth.Grid(16, 24)
    rr = synth.ribbon_roads(g, 400, seed=5)
    # crowd the roads' pairs towards the first tile rows: equal tile counts are then far from equal work
    keep = (rr.pairs.pair_tile < g.n_tiles // 3) | (np.arange(rr.pairs.n_pairs) % 4 == 0)
    road_of = rr.pairs.road_of_pair()[keep]
    from proj_roadsurf_b200.geometry import PairList
    pairs = PairList.from_pairs(400, road_of, rr.pairs.pair_tile[keep])
    world = 4
    cuts = balanced_tile_ranges(pairs.pair_tile, g.n_tiles, world)
    assert cuts[0] == 0 and cuts[-1] == g.n_tiles and np.all(np.diff(cuts) >= 0)
    per = [int(((pairs.pair_tile >= cuts[r]) & (pairs.pair_tile < cuts[r + 1])).sum()) for r in range(world)]
    even = [int(((pairs.pair_tile >= c0) & (pairs.pair_tile < c1)).sum()) for c0, c1 in zip(tile_ranges(g.n_tiles, world)[:-1],
                                                                                        tile_ranges(g.n_tiles, world)[1:])]
    assert max(per) - min(per) < 0.1 * pairs.n_pairs / world + 16 < max(even) - min(even)
    shards = plan_shards(None, pairs, g.n_tiles, world, balance="pairs")
    assert [s.tile_lo for s in shards] == list(cuts[:-1]) and sum(s.pairs.n_pairs for s in shards) == pairs.n_pairs
