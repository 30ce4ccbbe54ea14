#!/usr/bin/env python
"""bench.py — MLUPS of the D3Q27 hot path on the synthetic uniform box (BASELINE.json configs[1]).

  python bench.py --gpus N --steps K --warmup W            (ours: libludwig_b200.so through its C ABI)
  python bench.py --impl reference --gpus N --steps K --warmup W   (restated reference CPU path, host cores)

A "step" is one coarse time step of the whole hot path (K1 on every block of the level; the synthetic box has
no Bouzidi cells or refinement) over a 512^3 single-level box with open x faces and periodic y/z — the
configuration BASELINE.json's metric is quoted on.  Multi-GPU (N>1): one process per GPU, weak scaling: the box grows
to (512 N) x 512 x 512 cells, the library partitions its blocks along a Morton curve and K1 pulls the halo blocks of
other GPUs through NVLink peer mappings.  See DESIGN.md "Multi-GPU".

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BYTES_PER_LU = 216  # 27 x 4 B read + 27 x 4 B write
# dram__bytes_read.sum + dram__bytes_write.sum of k1_fast_kernel<PLAIN> per lattice update, from the ncu capture
# profiles/r1c_dram_traffic_k1_256cube.csv: (2 090 983 680 + 1 903 583 488) B / (30 720 blocks x 512 cells)
TRAFFIC_BYTES_PER_LU = (2090983680 + 1903583488) / (30720 * 512)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons during the run (NVML; falls back to nvidia-smi)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        flags = [bool(r & 0x8), bool(r & 0x40), bool(r & 0x20), bool(r & 0x4)]   # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
        util = n.nvmlDeviceGetUtilizationRates(self.handle).gpu
        return [str(sm), str(mx)] + ["Active" if f else "Not Active" for f in flags] + [util]

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.samples.append([s.strip() for s in out.split(",")] + [100])
            except Exception:
                pass
            self._stop_evt.wait(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        busy = [s for s in self.samples if s[6] and s[6] > 0] or self.samples      # samples taken under load
        sm = sorted(int(s[0]) for s in busy if str(s[0]).isdigit())
        mx = [int(s[1]) for s in self.samples if str(s[1]).isdigit()]
        reasons = set()
        for s in busy:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if str(v).lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples), "samples_under_load": len(busy)}


def make_params(cabi, n_cells_axis, strict):
    return cabi.Params(c_wale=0.5, nu_sgs_bg=0.0005, inlet_turbulence=0.01, q_min_threshold=0.001, wall_model_active=0,
                       use_temporal=0, sponge_blend=1, symmetric=0, domain_nx=n_cells_axis, domain_ny=n_cells_axis,
                       domain_nz=n_cells_axis, strict_fp=strict)


def cpu_leg(nb: int, steps: int, warmup: int):
    """Times the restated reference CPU path (oracle/, C++/OpenMP, -ffp-contract=off) on an nb^3-block box of
    the same recipe.  Returns (mlups, cores, sample description, ms_per_step)."""
    from open_ludwig_b200 import cabi
    from open_ludwig_b200.host import synthetic as syn
    lib = os.path.join(ROOT, "oracle", "_build", "libludwig_oracle.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    cores = os.cpu_count() or 1
    lv = syn.make_box_level(nb, nb, nb)
    f, rho, vel = syn.noise_state(lv)
    p = make_params(cabi, nb * 8, 1)
    with cabi.Context(lib) as c:
        c.add_level(lv)
        for w, a in ((cabi.F, f), (cabi.F_TEMP, f), (cabi.VEL, vel), (cabi.VEL_TEMP, vel), (cabi.RHO, rho)):
            c.upload(0, w, a)
        c.step_batch(1, warmup, 0.03, p)
        t0 = time.perf_counter()
        c.step_batch(1 + warmup, steps, 0.03, p)
        dt = time.perf_counter() - t0
    mlups = lv.n_cells * steps / dt / 1e6
    return mlups, cores, f"{nb * 8}^3 box ({lv.n_cells / 1e6:.2f} M cells), same recipe, {steps} steps, OpenMP {cores} threads", dt / steps * 1e3


def run_reference(args, rank, world):
    if rank != 0:
        return
    nb = args.cpu_nb
    mlups, cores, sample, ms = cpu_leg(nb, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "MLUPS (D3Q27 FP32)", "value": mlups, "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"synthetic uniform D3Q27 box, single level, inlet/outlet x + periodic y/z (bounded sample {nb * 8}^3 of the 512^3 workload)"},
        "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "restated reference CPU path (C++/OpenMP oracle, -ffp-contract=off); Julia/KernelAbstractions is not installed in this image",
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, local_rank, world):
    import copy
    import torch
    import torch.distributed as dist
    from open_ludwig_b200 import cabi
    from open_ludwig_b200 import multigpu as mg
    from open_ludwig_b200.host import synthetic as syn

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # weak scaling: 512^3 cells (64^3 blocks) per GPU; the global box is (64 N) x 64 x 64 blocks and the library's
    # Morton-range partition gives every rank one 64^3 cube (x is the most significant Morton digit here).
    nb = args.nb
    ncell_axis = nb * 8
    lv = syn.make_box_level(nb * world, nb, nb)
    p = cabi.Params(c_wale=0.5, nu_sgs_bg=0.0005, inlet_turbulence=0.01, q_min_threshold=0.001, wall_model_active=0, use_temporal=0,
                    sponge_blend=1, symmetric=0, domain_nx=ncell_axis * world, domain_ny=ncell_axis, domain_nz=ncell_axis, strict_fp=args.strict)
    ctx = cabi.Context(device=local_rank)
    if world > 1:
        ctx.set_partition(rank, world)
    ctx.add_level(lv)
    if world > 1:
        mg.attach_peers(ctx, dev)
    # initial state: equilibrium of a hashed (rho,u) field, generated for this rank's blocks only
    t0 = time.time()
    loc = ctx.local_blocks(0)
    if args.fast_init:
        # profiling runs only (ncu multiplies the cost of the 34 GB host-generated upload): rest state f = w_k on the device
        ctx.init_equilibrium()
    else:
        mine = copy.copy(lv)
        mine.active_block_coords = lv.active_block_coords[loc]
        f, rho, vel = syn.noise_state(mine)
        ctx.upload_local(0, cabi.F, f); ctx.upload_local(0, cabi.F_TEMP, f)
        ctx.upload_local(0, cabi.VEL, vel); ctx.upload_local(0, cabi.VEL_TEMP, vel); ctx.upload_local(0, cabi.RHO, rho)
        del f, rho, vel
    setup_s = time.time() - t0
    cells_per_rank = len(loc) * 512

    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank); sampler.start()
    t = 1
    barrier()
    ctx.step_batch(t, args.warmup, 0.03, p); t += args.warmup
    ctx.sync()

    # ---- timed region 1: device-resident throughput (value) + per-kernel timing of the dominant kernel
    ctx.profile_enable(True)
    n0 = ctx.launch_count()
    barrier()
    ev0.record(stream)
    ctx.step_batch(t, args.steps, 0.03, p); t += args.steps
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - n0
    k_ms, k_launches, k_cells = ctx.profile_read()
    ctx.profile_enable(False)

    # ---- timed region 2: end to end through the C ABI with host buffers: every step passes the host-side
    # params/u_inlet (kernel arguments) and reads the step's flow statistics back to the host.
    barrier()
    t0 = time.perf_counter()
    stats = None
    for _ in range(args.steps):
        ctx.step_batch(t, 1, 0.03, p); t += 1
        stats = ctx.flow_stats(0)          # device reduction + D2H of the per-CTA partials + host reduction (syncs)
        if world > 1:
            stats = mg.reduce_stats(stats, dev)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

    # ---- timed region 3: the same steps in the STRICT build (reference operation order, no FMA: bit-exact against the
    # CPU oracle) — reported beside the fast-mode headline, not instead of it
    strict_ms = None
    if world == 1 and not args.strict:
        ps = cabi.Params.from_buffer_copy(bytes(p)); ps.strict_fp = 1
        ctx.step_batch(t, 2, 0.03, ps); t += 2
        barrier()
        ev0.record(stream)
        ctx.step_batch(t, args.steps, 0.03, ps); t += args.steps
        ev1.record(stream)
        barrier()
        strict_ms = ev0.elapsed_time(ev1)

    times = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    ncell = torch.tensor([cells_per_rank], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(ncell)
    ms_max, e2e_ms_max = float(times[0]), float(times[1])
    total_cells = float(ncell[0])
    mlups = total_cells * args.steps / (ms_max * 1e-3) / 1e6
    e2e_mlups = total_cells * args.steps / (e2e_ms_max * 1e-3) / 1e6
    if rank == 0:
        peak, peak_src = measured_peaks()
        k_avg_ms = k_ms / max(k_launches, 1)
        achieved = (k_cells / max(k_launches, 1)) * BYTES_PER_LU / (k_avg_ms * 1e-3) / 1e9 if k_launches else None
        stats_parts = min(4096, max(1, min(148 * 8, (cells_per_rank + 255) // 256)))
        line = {
            "metric": "MLUPS (D3Q27 FP32)", "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"synthetic uniform {ncell_axis}^3 D3Q27 box per GPU ({ncell_axis * world}x{ncell_axis}x{ncell_axis} in total), single refinement "
                                   "level, inlet/outlet x + periodic y/z, regularized-BGK + WALE (c_wale 0.5, nu_bg 5e-4, inlet turbulence 0.01)",
                       "blocks_per_gpu": len(loc), "cells_per_gpu": cells_per_rank, "fp_mode": "strict" if args.strict else "fast",
                       "l2": f"working set {ctx.device_bytes() / 1e9:.1f} GB per GPU >> 126 MB L2, no flush needed",
                       "multi_gpu": ("Morton-range block partition; K1 pulls the remote halo layers over NVLink peer mappings (CUDA IPC) "
                                     "inside the stream-collide kernel; one stream-ordered peer-flag barrier kernel per step "
                                     "(no NCCL on the data path)") if world > 1 else "single GPU"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": TRAFFIC_BYTES_PER_LU * k_cells / max(k_launches, 1) if k_launches else None,
                         "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per LU (profiles/r1c_dram_traffic_k1_256cube.csv) x LU per launch",
                         "peak_source": peak_src, "kernel": "k1_fast_kernel<PLAIN> (rank 0)", "kernel_ms": k_avg_ms,
                         "bytes_per_lu": BYTES_PER_LU, "lu_per_launch": k_cells / max(k_launches, 1),
                         "frac_of_8TBs_nominal": (achieved / 8000.0) if achieved else None},
            "e2e": {"value": e2e_mlups, "unit": "MLUPS", "h2d_bytes_per_step": int(64), "d2h_bytes_per_step": int(stats_parts * 48),
                    "what": "ludwig_step_batch(1 step, host params) + ludwig_flow_stats (device reduction, D2H, host sync) every step"},
            "strict_mode": ({"value": total_cells * args.steps / (strict_ms * 1e-3) / 1e6, "unit": "MLUPS", "ms_per_step": strict_ms / args.steps,
                             "what": "same workload with strict_fp = 1: the reference's FP32 operation order without FMA contraction, "
                                     "bit-exact against the CPU oracle (tests/test_large_sizes_gpu.py)"} if strict_ms else None),
            "gpu_launches": int(launches),
            **({"not_a_bench_value": "--fast-init: rest-state initial condition, profiling run"} if args.fast_init else {}),
            "clocks": clocks,
            "setup_s": setup_s,
            "flow_stats_last": stats,
        }
        if world == 1 and not args.no_cpu:
            c_mlups, cores, sample, _ = cpu_leg(args.cpu_nb, args.cpu_steps, 1)
            line["cpu_baseline"] = {"value": c_mlups, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nb", type=int, default=64, help="blocks per axis (64 -> 512^3 cells)")
    ap.add_argument("--strict", type=int, default=0, help="1 = parity build (reference operation order, no FMA)")
    ap.add_argument("--cpu-nb", type=int, default=16, help="blocks per axis of the bounded CPU sample (16 -> 128^3)")
    ap.add_argument("--cpu-steps", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--fast-init", action="store_true", help="device-side rest-state initialisation instead of the hashed noise "
                    "state of config 2 (for ncu captures; the line is marked and is not a bench value)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
