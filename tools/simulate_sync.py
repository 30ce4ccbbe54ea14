"""What would finer-grained cross-rank synchronisation buy?  (CPU only: planning tool, no GPU.)

Input: the per-rank, per-level device times of `tools/emulate_ranks.py` (each virtual rank measured alone on one GPU)
and the actual partition of the case (spatially aligned plan).  A discrete-event simulation replays the sub-cycling
recursion of every rank with those durations under two synchronisation rules:

  barrier : a global barrier after every level step (what the library does);
  flags   : per-level completion counters, a rank waits only for the ranks it actually exchanges data with —
            same-level halo neighbours (RAW on their f_out / WAR on the A-B buffers), the owners of remote parent cells
            of its interface ghost cells, and the ranks whose children read its parent buffers.

    python tools/simulate_sync.py bunny_fine 8 profiles/r1e_emulate_8ranks_bunny_fine.log
"""
import ctypes as C, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from open_ludwig_b200 import cabi, partition
from open_ludwig_b200.host import domain as D
from open_ludwig_b200.host.cases import CASE_OVERRIDES, case_dir

name, world, log = sys.argv[1], int(sys.argv[2]), sys.argv[3]
T = {}
for line in open(log):
    m = re.match(r"rank (\d+): .*per level \[ms/coarse step\]: (.*)", line)
    if m:
        T[int(m.group(1))] = [float(x) for x in re.findall(r"L\d+ ([0-9.]+) \(", m.group(2))]
assert sorted(T) == list(range(world)), "need one 'rank r:' line per rank"
nl = len(T[0])
T = np.array([T[r] for r in range(world)])                  # [rank][level] ms per coarse step
dur = T / (2.0 ** np.arange(nl))[None, :]                   # ms per level step

case, ov = CASE_OVERRIDES[name]
dom = D.load_case(case_dir(case), ov, verbose=False, build_tri_map=False)
assert len(dom.levels) == nl
lib = cabi.load_library()
descs, keeps = zip(*[cabi.Context.make_desc(lv) for lv in dom.levels])
arr = (C.POINTER(cabi.LevelDesc) * nl)(*[C.pointer(d) for d in descs])
keys = (C.c_uint64 * (world + 1))()
assert lib.ludwig_partition_plan(arr, nl, world, keys) == 0
keys = np.array(list(keys), np.uint64)

# ownership of every block under the plan (abi.cu ludwig_level_create, has_plan branch)
owner = []
for l, lv in enumerate(dom.levels):
    order = partition.morton_order(lv.active_block_coords)
    c = np.asarray(lv.active_block_coords, np.int64)[order] - 1
    key = (partition._spread3(c[:, 0]) | (partition._spread3(c[:, 1]) << np.uint64(1)) | (partition._spread3(c[:, 2]) << np.uint64(2)))
    sk = key << np.uint64(3 * (nl - 1 - l))
    nb = len(order)
    st = [0]
    for r in range(1, world):
        cut = int(np.searchsorted(sk, keys[r], side="left"))
        cut = min(max(cut, st[-1] + 1), nb - (world - r))
        st.append(cut)
    st.append(nb)
    own_int = np.empty(nb, np.int32)
    for r in range(world):
        own_int[st[r]:st[r + 1]] = r
    own = np.empty(nb, np.int32); own[order] = own_int         # reference order
    owner.append(own)
    print(f"level {l+1}: blocks per rank {np.bincount(own, minlength=world).tolist()}", flush=True)

# who exchanges with whom
NB = [[set() for _ in range(world)] for _ in range(nl)]         # same-level halo neighbours
PB = [[set() for _ in range(world)] for _ in range(nl)]         # owners of (possibly) remote parents of level l blocks
for l, lv in enumerate(dom.levels):
    nt = np.asarray(lv.neighbor_table)                          # [27, nb] 1-based
    own = owner[l]
    for d in range(27):
        has = nt[d] > 0
        a, b = own[has], own[nt[d][has] - 1]
        for x, y in set(zip(a[a != b].tolist(), b[a != b].tolist())):
            NB[l][x].add(y); NB[l][y].add(x)
    if l > 0:
        # parents of a block AND of its 26 neighbour positions (ghost blocks interpolate there): parent block = (b-1)//2 (+-1 margin)
        pl = dom.levels[l - 1]
        bp = np.asarray(pl.block_pointer)                       # [bz,by,bx] 1-based
        c = np.asarray(lv.active_block_coords, np.int64) - 1
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    q = (c + np.array([dx, dy, dz])) // 2
                    ok = (q >= 0).all(axis=1) & (q[:, 0] < bp.shape[2]) & (q[:, 1] < bp.shape[1]) & (q[:, 2] < bp.shape[0])
                    pb = np.zeros(len(c), np.int64); pb[ok] = bp[q[ok, 2], q[ok, 1], q[ok, 0]]
                    sel = pb > 0
                    a, b = own[sel], owner[l - 1][pb[sel] - 1]
                    for x, y in set(zip(a[a != b].tolist(), b[a != b].tolist())):
                        PB[l][x].add(y)
for l in range(nl):
    print(f"level {l+1}: halo-neighbour ranks {[sorted(s) for s in NB[l]]}  remote-parent owners {[sorted(s) for s in PB[l]]}", flush=True)


def steps_in_order(n_coarse):
    """the recursion of solver_control.jl: (level, count) in issue order, count = 1-based step number of that level"""
    out, cnt = [], [0] * nl
    def rec(l):
        cnt[l] += 1; out.append((l, cnt[l]))
        if l + 1 < nl:
            rec(l + 1); rec(l + 1)
    for _ in range(n_coarse):
        rec(0)
    return out


def simulate(rule, n_coarse=6):
    seq = steps_in_order(n_coarse)
    end = {}                                                    # (rank, level, count) -> finish time
    t_rank = np.zeros(world)
    for (l, s) in seq:                                          # every rank issues the same order; process step by step
        start = t_rank.copy()
        if rule == "barrier":
            pass
        else:
            for r in range(world):
                deps = [end.get((q, l, s - 1), 0.0) for q in NB[l][r]]
                if l > 0:
                    p = (s + 1) // 2                            # parent step this child step interpolates from
                    deps += [end.get((q, l - 1, p), 0.0) for q in PB[l][r]]
                if l + 1 < nl:                                  # WAR: children (of any rank reading my buffers) of my previous step
                    deps += [end.get((q, l + 1, 2 * (s - 1)), 0.0) for q in range(world) if r in PB[l + 1][q]]
                if deps:
                    start[r] = max(start[r], max(deps))
        fin = start + dur[:, l]
        if rule == "barrier":
            fin[:] = fin.max()
        for r in range(world):
            end[(r, l, s)] = fin[r]
        t_rank = fin
    # steady-state period: time between the ends of the last two coarse steps
    per = []
    for k in range(n_coarse - 2, n_coarse):
        per.append(max(end[(r, nl - 1, (k + 1) * 2 ** (nl - 1))] for r in range(world)))
    return per[1] - per[0]


ideal = T.sum(axis=1).mean()
for rule in ("barrier", "flags"):
    print(f"SIMULATED {rule:8s}: {simulate(rule):.2f} ms per coarse step   (mean rank {ideal:.2f}, slowest rank {T.sum(axis=1).max():.2f}, "
          f"sum of level max {T.max(axis=0).sum():.2f})", flush=True)
