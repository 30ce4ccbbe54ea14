"""N > 1 host logic on CPU: two gloo ranks (world_size 2) run the partition rule the library uses, agree on a disjoint
cover, on symmetric halo ownership, and reduce per-rank flow statistics / partial forces exactly like bench.py and
tools/mg_check.py do on the GPUs (open_ludwig_b200/multigpu.py).  The GPU side of the same path (bit-identical to the
single-GPU run) is tools/mg_check.py, run with `gpurun --gpus 2`.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from open_ludwig_b200 import cabi, partition
from open_ludwig_b200.host import synthetic as syn
from open_ludwig_b200.host.cases import have_case

PORT = 29631


def test_partition_rule_matches_library(cuda_lib, oracle_lib):
    """ludwig_partition_starts is host code: callable without a GPU, identical in both libraries and in the mirror."""
    for path in (cuda_lib, oracle_lib):
        lib = C.CDLL(path)
        for n, w in ((96, 2), (1728, 8), (7, 3), (262144, 8)):
            out = (C.c_int32 * (w + 1))()
            assert lib.ludwig_partition_starts(n, w, out) == 0
            assert list(out) == list(partition.partition_starts(n, w))
        assert lib.ludwig_partition_starts(10, 9, None) != 0          # more than 8 ranks: rejected


def test_bench_box_partition_gives_cubes():
    """bench.py's weak-scaling box (64 N x 64 x 64 blocks): every rank's Morton range is one 64^3 cube (checked at 1/8 scale:
    8 N x 8 x 8).  The feature-less inlet / outlet blocks of level 1 ride in the plain launch with a body as lean as the plain one
    (merge_face), so they cost what a plain block costs and the cost-weighted cut gives every rank exactly one cube; on a walled box
    the blocks on the y / z faces keep the surcharge of the general domain-face class."""
    for world in (2, 4, 8):
        lv = syn.make_box_level(8 * world, 8, 8)
        cost = partition.block_costs(lv)
        assert cost.min() == 1.0 and cost.max() == 1.0                                         # periodic y / z: only x-only face blocks
        for r in range(world):
            for level in (None, lv):
                c = lv.active_block_coords[partition.local_blocks(lv.active_block_coords, r, world, level=level)]
                assert len(c) == 512 and c[:, 0].min() == 8 * r + 1 and c[:, 0].max() == 8 * r + 8
    walled = syn.make_box_level(8, 4, 4, periodic_y=False, periodic_z=False)
    cost = partition.block_costs(walled)
    co = walled.active_block_coords
    on_yz_face = (co[:, 1] == 1) | (co[:, 1] == 4) | (co[:, 2] == 1) | (co[:, 2] == 4)
    assert np.all(cost[on_yz_face] == 2.0) and np.all(cost[~on_yz_face] == 1.0)               # inlet / outlet interior: lean, 1.0


def test_block_costs_match_library(cuda_lib):
    """ludwig_block_costs is host code (no GPU needed): the NumPy mirror must agree on a case with every feature."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import test_k1_features_gpu as T
    lib = cabi.load_library(cuda_lib)
    for lv in T.build_case():
        nb = lv.n_blocks
        keep = [np.ascontiguousarray(a) for a in (lv.neighbor_table, lv.obstacle, lv.sponge, lv.wall_dist)]
        d = cabi.LevelDesc()
        d.n_blocks = nb
        d.level_id = lv.level_id
        d.neighbor_table, d.obstacle, d.sponge, d.wall_dist = (a.ctypes.data_as(C.c_void_p) for a in keep)
        d.bouzidi_enabled = int(lv.bouzidi_enabled)
        d.n_boundary_cells = lv.n_boundary_cells
        cb = np.ascontiguousarray(lv.cell_block, np.int32) if lv.cell_block is not None else None
        d.cell_block = cb.ctypes.data_as(C.c_void_p) if cb is not None else None
        out = np.zeros(nb, np.float32)
        assert lib.ludwig_block_costs(C.byref(d), out.ctypes.data_as(C.c_void_p)) == 0
        assert np.array_equal(out, partition.block_costs(lv))


def _worker(rank, world, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(PORT))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from open_ludwig_b200 import multigpu as mg
        import sys
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
        import test_k1_features_gpu as T
        levels = T.build_case()
        dev = torch.device("cpu")
        for lv in levels:
            mine = partition.local_blocks(lv.active_block_coords, rank, world, level=lv)
            # 1. disjoint cover, same order on every rank
            sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(sizes, torch.tensor([len(mine)]))
            pad = int(max(s.item() for s in sizes))
            buf = torch.full((pad,), -1, dtype=torch.int64); buf[:len(mine)] = torch.from_numpy(mine.astype(np.int64))
            allb = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(allb, buf)
            cat = np.concatenate([b.numpy()[:int(s.item())] for b, s in zip(allb, sizes)])
            assert sorted(cat.tolist()) == list(range(lv.n_blocks))
            assert np.array_equal(cat, partition.internal_order(lv.active_block_coords, world, level=lv))
            assert np.array_equal(np.sort(cat), np.sort(partition.morton_order(lv.active_block_coords)))
            # 2. halo symmetry: every remote block I pull from is owned by the peer, and the peer pulls from me too
            rem = partition.remote_neighbours(lv.neighbor_table, lv.active_block_coords, rank, world, level=lv)
            own = partition.owner_of_ref(lv.active_block_coords, world, level=lv)
            assert np.all(own[rem] != rank) and len(rem) > 0
            cnt = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(cnt, torch.tensor([len(rem)]))
            assert all(int(c.item()) > 0 for c in cnt)
        # 3. reductions: per-rank oracle statistics of the rank's own blocks combine to the statistics of the level
        lib = tmp["oracle"]
        lv = syn.make_box_level(4, 2, 2)
        f, rho, vel = syn.noise_state(lv)
        with cabi.Context(lib) as c:
            c.add_level(lv); c.upload(0, cabi.RHO, rho); c.upload(0, cabi.VEL, vel)
            whole = c.flow_stats(0)
        mine = partition.local_blocks(lv.active_block_coords, rank, world)
        sub = syn.make_box_level(len(mine), 1, 1)              # any level with the same number of blocks
        with cabi.Context(lib) as c:
            c.add_level(sub); c.upload(0, cabi.RHO, rho[mine]); c.upload(0, cabi.VEL, vel[:, mine])
            part = c.flow_stats(0)
        red = mg.reduce_stats(part, dev)
        assert red["n_fluid"] == whole["n_fluid"] and red["rho_min"] == whole["rho_min"] and red["rho_max"] == whole["rho_max"]
        assert red["v_max"] == whole["v_max"]
        assert abs(red["rho_mean"] - whole["rho_mean"]) < 1e-12 and abs(red["kinetic_energy"] - whole["kinetic_energy"]) < 1e-9
        aero = mg.reduce_aero({"Fx": 1.0 + rank, "Cd": 0.25 * (rank + 1)}, dev)
        assert aero["Fx"] == sum(1.0 + r for r in range(world)) and aero["Cd"] == 0.25 * sum(r + 1 for r in range(world))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_plumbing(oracle_lib):
    mp.spawn(_worker, args=(2, {"oracle": oracle_lib}), nprocs=2, join=True)


def _remote_pairs(lv, own, world):
    """(block, direction) pairs whose neighbour belongs to another rank, per rank — the halo surface of a partition"""
    nt = np.asarray(lv.neighbor_table)
    faces = np.zeros(world, np.int64)
    for d in range(27):
        has = nt[d] > 0
        a, b = own[has], own[nt[d][has] - 1]
        np.add.at(faces, a[a != b], 1)
    return faces


@pytest.mark.skipif(not have_case("ball1m"), reason="case folder not available")
def test_rcb_partition_is_balanced_compact_and_deterministic(cuda_lib):
    """ludwig_partition_rcb (host code, no GPU): equal cost per rank, every rank non-empty, same answer twice, and a smaller
    halo surface than the Morton-range cut on a real refined level (shipped ball1m, finest level: a shell around the sphere)."""
    import ctypes as C
    from open_ludwig_b200 import cabi
    from open_ludwig_b200.host import domain as D
    from open_ludwig_b200.host.cases import CASE_OVERRIDES, case_dir
    lib = cabi.load_library(cuda_lib)
    case, ov = CASE_OVERRIDES["sphere_re10m"]
    dom = D.load_case(case_dir(case), ov, build_tri_map=False)
    for world in (2, 3, 8):
        for lv in dom.levels[1:]:
            d, keep = cabi.Context.make_desc(lv)
            own = np.full(lv.n_blocks, -1, np.int32); own2 = own.copy()
            assert lib.ludwig_partition_rcb(C.byref(d), world, own.ctypes.data_as(C.c_void_p)) == 0
            assert lib.ludwig_partition_rcb(C.byref(d), world, own2.ctypes.data_as(C.c_void_p)) == 0
            assert np.array_equal(own, own2) and own.min() == 0 and own.max() == world - 1
            cost = partition.block_costs(lv).astype(np.float64)
            per = np.bincount(own, weights=cost, minlength=world)
            assert per.min() > 0 and per.max() - per.min() <= 0.02 * per.mean() + 2 * cost.max(), (world, lv.level_id, per)
    lv = dom.levels[-1]
    d, keep = cabi.Context.make_desc(lv)
    own = np.empty(lv.n_blocks, np.int32)
    assert lib.ludwig_partition_rcb(C.byref(d), 8, own.ctypes.data_as(C.c_void_p)) == 0
    morton = partition.owner_of_ref(lv.active_block_coords, 8, level=lv)
    assert _remote_pairs(lv, own, 8).max() < _remote_pairs(lv, morton, 8).max()
    own_bad = np.empty(3, np.int32)
    small = syn.make_box_level(1, 1, 2)
    d2, keep2 = cabi.Context.make_desc(small)
    assert lib.ludwig_partition_rcb(C.byref(d2), 4, own_bad.ctypes.data_as(C.c_void_p)) != 0      # fewer blocks than ranks


def test_xslab_order_keeps_x_neighbours_close():
    """The library's default internal order (x-slab, T = 12: DESIGN.md section 4): inside a rank every block's x neighbour is at most
    T * T positions away (its x-face halo sectors — 8x as expensive as a y / z face — are still in L2 when it runs), the order is a
    permutation, and it is the Morton order when block_order = 0."""
    lv = syn.make_box_level(20, 30, 26)
    co = lv.active_block_coords
    for T in (2, 5, 12):
        order = partition.internal_order(co, 1, block_order=T)
        assert sorted(order.tolist()) == list(range(lv.n_blocks))
        pos = np.empty(lv.n_blocks, np.int64); pos[order] = np.arange(lv.n_blocks)
        index = {tuple(c): i for i, c in enumerate(co.tolist())}
        worst = 0
        for i, (x, y, z) in enumerate(co.tolist()):
            j = index.get((x + 1, y, z))
            if j is not None:
                worst = max(worst, abs(int(pos[j]) - int(pos[i])))
        assert worst <= T * T, (T, worst)
    assert np.array_equal(partition.internal_order(co, 1, block_order=0), partition.morton_order(co))
    # partitioned: the owners are cut on the Morton order, only the order INSIDE a rank changes
    for world in (2, 3):
        own_m = partition.owner_of_ref(co, world, level=lv)
        for r in range(world):
            mine = partition.local_blocks(co, r, world, level=lv)
            assert np.all(own_m[mine] == r) and len(mine) == int((own_m == r).sum())
