"""Strong scaling of a named case over N GPUs (one process per GPU; launch with torchrun, or plain python for N = 1).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/run_case_mg.py bunny_fine 6 [--fp-mode strict] [--ramp 16] [--profile 2] \
        [--variant "partition=rcb_yz"] [--variant "plan"] [--variant "partition=rcb,fork_max_blocks=100000"] ...

The domain is built ONCE (rank 0, every host thread; the other ranks map it from /dev/shm) and every --variant — a comma-
separated list of library options (ludwig_ctx_set_option), "fp=fast|strict" and/or the word "plan" (spatially aligned partition plan) — is
measured on the same processes, one after the other: contexts are re-created, the domain is not.  Prints one RESULT line per
variant (true MLUPS from the max over ranks of CUDA-event device time) and, with --profile, every rank's per-level,
per-kernel-class device times.
"""
import argparse
import json
import os
import sys

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
# host threads (set before anything loads an OpenMP runtime; torchrun exports OMP_NUM_THREADS=1): rank 0 builds the domain
# with every core, the other ranks only map the cached arrays
os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 8) if rank == 0 else "2"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist   # noqa: E402
from open_ludwig_b200 import multigpu as mg   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("case"); ap.add_argument("steps", type=int)
ap.add_argument("--fp-mode", default="fast", choices=["fast", "strict"])
ap.add_argument("--ramp", type=int, default=None, help="ramp length override (coarse steps)")
ap.add_argument("--profile", type=int, default=0, help="coarse steps of the per-level profiling pass")
ap.add_argument("--uniform-start", action="store_true", help="start from a uniform flow at u_target instead of rest + ramp (O(1) forces at once)")
ap.add_argument("--variant", action="append", default=[])
ap.add_argument("--json", default=None, help="write the records here (rank 0)")
args = ap.parse_args()

torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
log = lambda m: print(m, flush=True)
dom, build_s = mg.load_domain_shared(args.case, log)
records = []
for v in (args.variant or [""]):
    items = [x for x in v.split(",") if x]
    plan = "plan" in items
    opts = dict(x.split("=", 1) for x in items if x != "plan")
    pov = {k[2:]: float(opts.pop(k)) for k in [k for k in opts if k.startswith("p.")]}   # "p.wall_model_active=0": ludwig_params override (experiments)
    fp = opts.pop("fp", args.fp_mode)                      # "fp=fast" / "fp=strict" inside a variant overrides --fp-mode
    rec = mg.run_case_strong(args.case, args.steps, lr, strict=fp == "strict", options=opts, plan=plan, ramp_steps=args.ramp,
                             profile_steps=args.profile, log=log, dom=dom, all_ranks_levels=args.profile > 0, uniform_start=args.uniform_start, param_overrides=pov)
    rec["variant"] = v or "default"; rec["domain_build_s"] = build_s
    records.append(rec)
    if rank == 0:
        print(f"RESULT case={args.case} variant=[{rec['variant']}] n_gpus={world} fp={rec['fp_mode']} steps={args.steps} ms/coarse_step={rec['ms_per_coarse_step']:.2f} "
              f"true_MLUPS={rec['mlups_true']:.0f} ref_MLUPS={rec['mlups_reference_style']:.0f} Cd={rec['Cd']:.6e} Cl={rec['Cl']:.6e} "
              f"rho_min={rec['rho_min']:.6f} rho_max={rec['rho_max']:.6f} upload_s={rec['upload_s']:.1f}", flush=True)
        per_rank = rec["all_ranks"] or ([{"blocks": rec["blocks_rank0"], "levels_ms": rec["rank0_levels_ms"]}] if rec["rank0_levels_ms"] else [])
        for r, d in enumerate(per_rank):
            print(f"  rank {r}: blocks {d['blocks']} per level [ms/coarse step]: " + " | ".join(
                f"L{i+1} {x['level_step']:.2f} (k1p {x['k1_plain']:.2f} pg {x['k1_plain_ghost']:.2f} ft {x['k1_feature']:.2f} fu {x['k1_full']:.2f} "
                f"pre {x['interface_prepass']:.2f} bz {x['bouzidi']:.2f} bar {x['barrier']:.2f})" for i, x in enumerate(d["levels_ms"])), flush=True)
if rank == 0 and args.json:
    with open(args.json, "w") as fh:
        json.dump(records, fh, indent=1)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
