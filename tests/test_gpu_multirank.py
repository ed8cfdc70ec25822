"""Two NCCL ranks on two GPUs (skipped with fewer): the sharded path -- fused kernel per rank, boundary rows merged by the C
ABI's rs_allreduce_accumulators_dev (one grouped NCCL launch), statistics finalized after the merge -- against the single-shot
result and the plain-C oracle.  -m gpu."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from proj_roadsurf_b200 import synth
    from proj_roadsurf_b200.distributed import global_rows, merge_boundary, plan_shards
    from proj_roadsurf_b200.engine import Engine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    g = synth.Grid(12, 16)
    rr = synth.ribbon_roads(g, 200, seed=21)
    sh = plan_shards(rr.roads, rr.pairs, g.n_tiles, world, only_rank=rank, balance="pairs")[rank]
    idx = np.arange(sh.tile_lo, sh.tile_hi)
    eng = Engine(rank)
    eng.comm_init_from_torch()
    assert eng.comm_world == world
    tiles = eng.synth_tiles_dev(g.keys(idx), 256, 256, 3, kind=0, gt=g.transforms(idx))
    slot = torch.from_numpy(sh.slot).to(dev)
    mz = torch.zeros((sh.n_rows,), dtype=torch.int32, device=dev)
    hist, nz = eng.zonal_hist_dev(eng.upload_roads(sh.roads), tiles, eng.upload_pairs(sh.pairs), road_slot=slot, n_slots=sh.n_rows,
                                  min_zero=mz)
    assert hist.shape[0] == sh.n_rows                                   # also the boundary rows this rank does not touch
    merge_boundary(hist, nz, sh.n_own, engine=eng, min_zero=mz)
    st = eng.finalize_stats_dev(hist, nz, nodata_mode="none", ddof=1)
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), hist=hist.cpu().numpy(), nz=nz.cpu().numpy(), mz=mz.cpu().numpy(),
             stats=st.cpu().numpy(), rows=global_rows(sh), n_own=sh.n_own)
    eng.close()
    dist.destroy_process_group()


def test_two_rank_nccl_merge_equals_single_shot_and_oracle(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from oracle import cport
    from proj_roadsurf_b200 import synth
    from proj_roadsurf_b200.engine import Engine
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = synth.Grid(12, 16)
    rr = synth.ribbon_roads(g, 200, seed=21)
    eng = Engine(0)
    tiles = eng.synth_tiles_dev(g.keys(), 256, 256, 3, kind=0, gt=g.transforms())
    mz1 = torch.zeros((200,), dtype=torch.int32, device="cuda:0")
    h1, z1 = eng.zonal_hist_dev(eng.upload_roads(rr.roads), tiles, eng.upload_pairs(rr.pairs), min_zero=mz1)
    st1 = eng.finalize_stats_dev(h1, z1, nodata_mode="none", ddof=1).cpu().numpy()
    oh, onz, omz = cport.zonal_accumulate(rr.roads.xy, rr.roads.ring_off, rr.roads.road_ring_off, rr.pairs.road_pair_off,
                                          rr.pairs.pair_tile, tiles.pixels.cpu().numpy(), g.transforms(), want_min_zero=True)
    h1, z1, mz1 = (t.cpu().numpy().view(np.uint32).astype(np.uint64) for t in (h1, z1, mz1))
    assert np.array_equal(h1, oh) and np.array_equal(z1, onz) and np.array_equal(mz1, omz)
    n_b = 0
    for r in range(world):
        d = np.load(tmp_path / f"r{r}.npz")
        rows = d["rows"]
        assert np.array_equal(d["hist"].view(np.uint32).astype(np.uint64), oh[rows])
        assert np.array_equal(d["nz"].view(np.uint32).astype(np.uint64), onz[rows])
        assert np.array_equal(d["mz"].view(np.uint32).astype(np.uint64), omz[rows])
        assert np.array_equal(d["stats"], st1[rows], equal_nan=True)     # finalized after the merge: identical tables
        n_b = len(rows) - int(d["n_own"])
    assert n_b > 0
    eng.close()
