"""Per-rank cost of a partitioned case, measured on ONE GPU (no multi-GPU box time needed).

    python tools/emulate_ranks.py bunny_fine 8 2 [--no-plan] [--opt partition=rcb_yz] [--opt fork_max_blocks=100000]

All `world` rank contexts are created on device 0 from the same domain and attached in-process
(ludwig_attach_inprocess), with a no-op barrier; each virtual rank then steps ALONE while the others' state stays
static, so its "remote" neighbour blocks are ordinary local memory.  What this shows: the share of every rank per
level and per kernel class (load balance of the partition, small-grid and launch overheads).  What it cannot show:
NVLink latency of the halo pulls and barrier skew — those are the difference to the real N-GPU run.
With barriers after every level step the N-GPU step time is bounded below by  sum over levels of  max over ranks.
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from open_ludwig_b200 import cabi
from open_ludwig_b200.host import domain as D
from open_ludwig_b200.host.cases import CASE_OVERRIDES, case_dir
from open_ludwig_b200.solver import make_params, ramp_velocity

name, world, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
opts = dict(a.split("=", 1) for i, a in enumerate(sys.argv) if i > 0 and sys.argv[i - 1] == "--opt")
use_plan = "--no-plan" not in sys.argv and "partition" not in opts
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
case, ov = CASE_OVERRIDES[name]
t0 = time.time()
dom = D.load_case(case_dir(case), ov, verbose=False, build_tri_map=False)
nl = len(dom.levels)
print(f"domain build {time.time()-t0:.1f}s cells {dom.total_cells/1e6:.1f}M updates/coarse step {dom.cell_updates_per_coarse_step/1e6:.0f}M", flush=True)
params = make_params(dom, strict=False)
u = ramp_velocity(dom.cfg.u_target, 8, dom.cfg.ramp_steps)


def timed(ctx, n):
    """(total ms, per-level list of class dicts) of n coarse steps of one context."""
    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.step_batch(1, 1, u, params); ctx.sync()
    e0.record(stream); ctx.step_batch(3, n, u, params); e1.record(stream); ctx.sync()
    total = e0.elapsed_time(e1)
    ctx.profile_enable(True)
    ctx.step_batch(3, n, u, params); ctx.sync()
    ctx.profile_read(); lv = ctx.profile_levels()
    ctx.profile_enable(False)
    return total, lv


ctxs = []
for r in range(world):
    c = cabi.Context(device=0, options=opts)
    c.set_partition(r, world)
    if use_plan:
        c.set_partition_plan(dom.levels)
    for lv in dom.levels:
        c.add_level(lv)
    ctxs.append(c)
cabi.Context.attach_inprocess(ctxs)
for c in ctxs:
    c.set_barrier(lambda: None)
    c.init_equilibrium()
for c in ctxs:
    c.step_batch(1, 0, u, params)      # zero steps: builds the tables and exports every rank's halo layers once
    c.sync()
print(f"{world} virtual ranks on one GPU, device GB total {sum(c.device_bytes() for c in ctxs)/1e9:.1f}, plan={use_plan}", flush=True)
tot = np.zeros(world); lvl = np.zeros((world, nl))
for r, c in enumerate(ctxs):
    t, lv = timed(c, steps)
    tot[r] = t / steps
    lvl[r] = [d["level_step"] / steps for d in lv]
    nloc = [len(c.local_blocks(i)) for i in range(nl)]
    print(f"rank {r}: {tot[r]:.2f} ms/coarse step  blocks {nloc}  per level [ms/coarse step]: " +
          " | ".join(f"L{i+1} {d['level_step']/steps:.2f} (pre {d['interface_prepass']/steps:.2f} k1p {d['k1_plain']/steps:.2f} bz {d['bouzidi']/steps:.2f})" for i, d in enumerate(lv)), flush=True)
for c in ctxs:
    c.close()
torch.cuda.empty_cache()
one = cabi.Context(device=0)
for lv in dom.levels:
    one.add_level(lv)
one.init_equilibrium(); one.sync()
t1, lv1 = timed(one, steps)
t1 /= steps
print(f"1 rank: {t1:.2f} ms/coarse step  per level: " + " | ".join(f"L{i+1} {d['level_step']/steps:.2f} (pre {d['interface_prepass']/steps:.2f} k1p {d['k1_plain']/steps:.2f} bz {d['bouzidi']/steps:.2f})" for i, d in enumerate(lv1)), flush=True)
one.close()
bound = lvl.max(axis=0).sum()
print(f"EMULATE case={name} world={world} plan={use_plan} opts={opts} ideal={t1/world:.2f} mean_rank={tot.mean():.2f} max_rank={tot.max():.2f} "
      f"sum_of_level_max={bound:.2f} ms  -> efficiency bounds: overhead-only {t1/world/tot.mean():.3f}, +imbalance {t1/world/tot.max():.3f}, "
      f"+per-level barriers {t1/world/bound:.3f}", flush=True)
