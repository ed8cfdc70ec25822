"""Tile sharding over the GPUs of one box and the merge of per-road accumulators.

The path shards by tile: tiles are disjoint pixel sets and a road's statistics are a fold of
associative, commutative integer additions over its (road, tile) pairs
(scripts/statistical_analysis/statistical_analysis.py:187-193 concatenates a road's pixels over
all its tiles before the groupby).  Each rank owns a contiguous block of the tile index range and
every pair whose tile it owns.  Roads whose pairs all fall on one rank are final there.  Roads
that cross onto tiles of several ranks ("boundary roads", a few per cent for row-band shards) get a
row in a boundary table that has the same layout on every rank; one all-reduce(SUM) of that table
(uint32 counters; NCCL over NVLink on GPUs, gloo in the CPU tests) completes them everywhere.
Median / percentiles are computed after the merge, from merged histograms, hence exact.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .geometry import PairList, RoadSet


@dataclass
class Shard:
    rank: int
    world: int
    tile_lo: int                 # owned tile index range [tile_lo, tile_hi)
    tile_hi: int
    road_global: np.ndarray      # (n_local,) global index of each local road; own roads first, then boundary roads
    n_own: int                   # local roads final on this rank
    slot: np.ndarray             # (n_local,) int32 output row of each local road
    n_rows: int                  # rows of the output tables = n_own + n_boundary
    boundary_global: np.ndarray  # (n_boundary,) global road index of each boundary-table row (same on all ranks)
    roads: Optional[RoadSet] = None
    pairs: Optional[PairList] = None   # local roads x local tile indices (rebased to tile_lo)

    @property
    def n_boundary(self) -> int:
        return int(len(self.boundary_global))


def tile_ranges(n_tiles: int, world: int) -> np.ndarray:
    """(world + 1,) tile index cuts; equal contiguous blocks (row bands of a row-major lattice)."""
    return (np.arange(world + 1, dtype=np.int64) * n_tiles) // world


def balanced_tile_ranges(pair_tile: np.ndarray, n_tiles: int, world: int) -> np.ndarray:
    """(world + 1,) tile index cuts of contiguous blocks that hold (nearly) equal numbers of (road, tile) pairs: the
    kernel's work follows the pairs, not the tiles, so these cuts level the per-rank kernel times."""
    per_tile = np.bincount(np.asarray(pair_tile, np.int64), minlength=n_tiles)
    csum = np.concatenate([[0], np.cumsum(per_tile)])
    target = csum[-1] * np.arange(world + 1, dtype=np.float64) / world
    cuts = np.searchsorted(csum, target, side="left").astype(np.int64)
    cuts[0], cuts[-1] = 0, n_tiles
    return np.maximum.accumulate(np.minimum(cuts, n_tiles))


def plan_shards(roads: Optional[RoadSet], pairs: PairList, n_tiles: int, world: int,
                only_rank: Optional[int] = None, balance: str = "tiles") -> List[Shard]:
    """Split a global pair list by tile block.  With `roads`, each shard carries its compact road soup.
    balance: 'tiles' = equal tile counts per rank, 'pairs' = contiguous tile ranges with equal pair counts."""
    cuts = tile_ranges(n_tiles, world) if balance == "tiles" or world == 1 else balanced_tile_ranges(pairs.pair_tile, n_tiles, world)
    R = len(pairs.road_pair_off) - 1
    road_of = pairs.road_of_pair().astype(np.int64)
    pair_rank = np.searchsorted(cuts, pairs.pair_tile.astype(np.int64), side="right") - 1
    # per road: first and last rank touched (pairs are sorted by tile within a road, so ranks too)
    has = np.diff(pairs.road_pair_off) > 0
    rmin = np.full(R, world, np.int64)
    rmax = np.full(R, -1, np.int64)
    np.minimum.at(rmin, road_of, pair_rank)
    np.maximum.at(rmax, road_of, pair_rank)
    boundary = has & (rmax > rmin)
    boundary_global = np.nonzero(boundary)[0]
    b_index = np.full(R, -1, np.int64)
    b_index[boundary_global] = np.arange(len(boundary_global))
    shards = []
    for r in range(world):
        if only_rank is not None and r != only_rank:
            shards.append(None)
            continue
        lo, hi = int(cuts[r]), int(cuts[r + 1])
        touch = np.zeros(R, bool)
        touch[road_of[pair_rank == r]] = True
        own = np.nonzero(touch & ~boundary)[0]
        bnd = np.nonzero(touch & boundary)[0]
        road_global = np.concatenate([own, bnd])
        slot = np.concatenate([np.arange(len(own)), len(own) + b_index[bnd]]).astype(np.int32)
        local_pairs = pairs.restrict_tiles(lo, hi).take_roads(road_global)
        shards.append(Shard(r, world, lo, hi, road_global, int(len(own)), slot, int(len(own) + len(boundary_global)),
                            boundary_global, None if roads is None else roads.subset(road_global), local_pairs))
    return shards


def merge_boundary(hist, n_allzero, n_own: int, group=None, engine=None, min_zero=None):
    """In-place all-reduce(SUM) of the boundary rows of a rank's tables (torch tensors, int32 storage of
    the uint32 counters: two's-complement addition is the same bit pattern).  With an ``engine`` that holds a
    communicator (Engine.comm_init*) the merge is the C ABI's rs_allreduce_accumulators_dev (one grouped NCCL launch);
    otherwise torch.distributed collectives (gloo in the CPU tests)."""
    if engine is not None and engine.comm_world > 1:
        engine.allreduce_accumulators_dev(hist, n_allzero, n_own, min_zero)
        return
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    if hist.shape[0] > n_own:
        dist.all_reduce(hist[n_own:], op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(n_allzero[n_own:], op=dist.ReduceOp.SUM, group=group)
        if min_zero is not None:
            dist.all_reduce(min_zero[n_own:], op=dist.ReduceOp.SUM, group=group)


def global_rows(shard: Shard) -> np.ndarray:
    """(n_rows,) global road index of every output row of this shard (own rows, then the boundary
    table, whose rows are valid on every rank after the merge)."""
    g = np.empty(shard.n_rows, np.int64)
    g[:shard.n_own] = shard.road_global[:shard.n_own]
    g[shard.n_own:] = shard.boundary_global
    return g
