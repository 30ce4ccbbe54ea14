"""Partition quality of a case for N ranks, computed on the CPU (planning tool, no GPU): for every partition rule of the
library — cost-weighted Morton ranges per level ("morton"), the spatially aligned plan ("plan"), per-level recursive coordinate
bisection ("rcb") and RCB that never cuts across x ("rcb_yz") — per level:

  * load balance: max over ranks / mean of the estimated cost (ludwig_block_costs x 2^(level-1) sub-steps);
  * halo surface: remote (block, direction) pairs a rank pulls through, split by the orientation of the cut (x-face / y-face /
    z-face / edge+corner) and turned into NVLink bytes per level step with the sector amplification of the block layout
    (an x-face layer is 64 separate 32-byte sectors per direction for 4 useful bytes each: 8x; y / z faces: 1x);
  * remote parents: share of a fine level's interface ghost blocks whose parent block lives on another rank.

    python tools/partition_stats.py bunny_fine 8
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from open_ludwig_b200 import cabi, partition  # noqa: E402
from open_ludwig_b200.host import domain as D  # noqa: E402
from open_ludwig_b200.host.cases import CASE_OVERRIDES, case_dir  # noqa: E402

name, world = sys.argv[1], int(sys.argv[2])
case, ov = CASE_OVERRIDES[name]
dom = D.load_case(case_dir(case), ov, verbose=False, build_tri_map=False)
nl = len(dom.levels)
lib = cabi.load_library()
descs, keeps = zip(*[cabi.Context.make_desc(lv) for lv in dom.levels])
costs = []
for d, lv in zip(descs, dom.levels):
    c = np.empty(lv.n_blocks, np.float32)
    assert lib.ludwig_block_costs(C.byref(d), c.ctypes.data_as(C.c_void_p)) == 0
    costs.append(c)


def morton_keys(lv):
    c = np.asarray(lv.active_block_coords, np.int64) - 1
    return partition._spread3(c[:, 0]) | (partition._spread3(c[:, 1]) << np.uint64(1)) | (partition._spread3(c[:, 2]) << np.uint64(2))


def owners(rule):
    out = []
    if rule == "plan":
        arr = (C.POINTER(cabi.LevelDesc) * nl)(*[C.pointer(d) for d in descs])
        keys = (C.c_uint64 * (world + 1))()
        assert lib.ludwig_partition_plan(arr, nl, world, keys) == 0
        keys = np.array(list(keys), np.uint64)
    for l, lv in enumerate(dom.levels):
        nb = lv.n_blocks
        if rule in ("rcb", "rcb_yz"):
            own = np.empty(nb, np.int32)
            assert lib.ludwig_partition_rcb_axes(C.byref(descs[l]), world, 7 if rule == "rcb" else 6, own.ctypes.data_as(C.c_void_p)) == 0
        else:
            key = morton_keys(lv)
            order = np.argsort(key, kind="stable")
            if rule == "morton":
                st = partition.weighted_starts(costs[l][order], world)
            else:
                sk = key[order] << np.uint64(3 * (nl - 1 - l))
                st = [0]
                for r in range(1, world):
                    cut = int(np.searchsorted(sk, keys[r], side="left"))
                    st.append(min(max(cut, st[-1] + 1), nb - (world - r)))
                st.append(nb)
            own = np.empty(nb, np.int32)
            for r in range(world):
                own[order[st[r]:st[r + 1]]] = r
        out.append(own)
    return out


FACE_BYTES = {"x": 9 * 64 * 32 + 3 * 64 * 32, "y": 9 * 8 * 32 + 3 * 8 * 32, "z": 9 * 256 + 3 * 256}   # populations + velocities, sectors touched
EDGE_BYTES = 3 * 8 * 32       # upper bound: 8 cells x 3 directions, one sector each (x-parallel edges need 1 sector per direction)
for rule in ("plan", "morton", "rcb", "rcb_yz"):
    own = owners(rule)
    print(f"=== {name}, {world} ranks, partition = {rule}")
    tot_cost = np.zeros(world); tot_bytes = np.zeros(world)
    for l, lv in enumerate(dom.levels):
        sub = 2 ** l
        o = own[l]
        cost = np.bincount(o, weights=costs[l], minlength=world) * sub
        nt = np.asarray(lv.neighbor_table)                       # [27, nb] 1-based
        pairs = {"x": np.zeros(world), "y": np.zeros(world), "z": np.zeros(world), "e": np.zeros(world)}
        for d in range(27):
            if d == 13:
                continue
            dx, dy, dz = d % 3 - 1, (d // 3) % 3 - 1, d // 9 - 1
            nz = (dx != 0) + (dy != 0) + (dz != 0)
            kind = "e" if nz > 1 else ("x" if dx else "y" if dy else "z")
            v = nt[d]
            has = v > 0
            remote = has & (o[np.maximum(v - 1, 0)] != o)
            pairs[kind] += np.bincount(o[remote], minlength=world)
        byts = pairs["x"] * FACE_BYTES["x"] + pairs["y"] * FACE_BYTES["y"] + pairs["z"] * FACE_BYTES["z"] + pairs["e"] * EDGE_BYTES
        remote_parent = ""
        if l > 0:
            # parent block of every fine block: coords (b-1)//2 on level l-1
            pc = (np.asarray(lv.active_block_coords, np.int64) - 1) // 2
            P = dom.levels[l - 1]
            bp = np.asarray(P.block_pointer)                      # [dimz, dimy, dimx] 1-based
            ok = (pc[:, 0] < bp.shape[2]) & (pc[:, 1] < bp.shape[1]) & (pc[:, 2] < bp.shape[0])
            pidx = np.zeros(len(pc), np.int64)
            pidx[ok] = bp[pc[ok, 2], pc[ok, 1], pc[ok, 0]]
            has_p = pidx > 0
            iface = (nt == 0).any(axis=0)                          # blocks with a missing neighbour (interface or domain face)
            rp = has_p & iface & (own[l - 1][np.maximum(pidx - 1, 0)] != o)
            remote_parent = f" remote-parent interface blocks {rp.sum()}/{int((has_p & iface).sum())}"
        tot_cost += cost; tot_bytes += byts * sub
        print(f"  L{l + 1}: blocks/rank {np.bincount(o, minlength=world).tolist()}  cost max/mean {cost.max() / cost.mean():.3f}  "
              f"remote pairs x {int(pairs['x'].sum())} y {int(pairs['y'].sum())} z {int(pairs['z'].sum())} edge {int(pairs['e'].sum())}  "
              f"NVLink MB/level step: max rank {byts.max() / 1e6:.1f}, mean {byts.mean() / 1e6:.1f}{remote_parent}")
    print(f"  total: cost max/mean {tot_cost.max() / tot_cost.mean():.3f} (sum over levels of per-level max / mean total: "
          f"{sum((np.bincount(own[l], weights=costs[l], minlength=world) * 2 ** l).max() for l in range(nl)) / tot_cost.mean():.3f})  "
          f"NVLink GB per coarse step: max rank {tot_bytes.max() / 1e9:.2f}, mean {tot_bytes.mean() / 1e9:.2f}", flush=True)
