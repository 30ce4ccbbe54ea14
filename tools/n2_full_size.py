"""N2 at full size: the 339 M-cell bunny (config 5) built twice — host threads only, and with the three brute-force phases on the
device (ludwig_domain_voxelize / _wall_distance / _qmap) — phase times side by side and every table compared byte for byte."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from open_ludwig_b200.host import domain as D
from open_ludwig_b200.host.cases import CASE_OVERRIDES, case_dir

name = sys.argv[1] if len(sys.argv) > 1 else "bunny_fine"
case, ov = CASE_OVERRIDES[name]
t0 = time.time(); host = D.load_case(case_dir(case), ov, build_tri_map=False); t1 = time.time()
dev = D.load_case(case_dir(case), ov, build_tri_map=False, gpu_device=0); t2 = time.time()
print(f"{name}: {host.total_cells / 1e6:.1f} M cells, host threads {len(os.sched_getaffinity(0))}")
print(f"host build   {t1 - t0:6.1f} s  phases " + ", ".join(f"{k} {v:.2f}" for k, v in host.phase_s.items()))
print(f"device build {t2 - t1:6.1f} s  phases " + ", ".join(f"{k} {v:.2f}" for k, v in dev.phase_s.items()))
ok = True
for a, b in zip(host.levels, dev.levels):
    for k in ("obstacle", "wall_dist", "q_map", "cell_block", "cell_x", "cell_y", "cell_z"):
        x, y = getattr(a, k), getattr(b, k)
        same = (x is None and y is None) or (x is not None and y is not None and x.shape == y.shape and x.tobytes() == y.tobytes())
        ok &= same
        if not same:
            print("MISMATCH level", a.level_id, k)
print("N2_FULL_SIZE", "IDENTICAL" if ok else "DIFFERENT", "boundary cells", [lv.n_boundary_cells for lv in dev.levels])
