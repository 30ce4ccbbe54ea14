"""Where the reference's case folders live and the named configurations of SURVEY.md §8(d).

The case files (config.yaml + STL) are the reference's own input format, kept verbatim.  In the build container
they are read from /root/reference/CASES; ``tools/fetch_cases.py`` copies them to baseline/_ref/CASES (git-ignored,
shipped to the GPU box with the working tree) because /root/reference does not exist there.
"""
from __future__ import annotations

import os

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SEARCH = (os.path.join(_ROOT, "baseline", "_ref", "CASES"), "/root/reference/CASES", os.path.join(_ROOT, "cases"))

# name -> (case folder, overrides merged over its config.yaml)
CASE_OVERRIDES = {
    # config 3: the configuration RESULTS_SPHERE_RE1M.txt was produced with (N = 25 cells/L, U = 14.8 m/s; :38,:153)
    "sphere_re1m": ("ball1m", {"basic": {"surface_resolution": 25, "flow": {"velocity": 14.8}}}),
    # the same grid at U = 4.0 m/s: the configuration of RESULTS_SPHERE_RE266K.txt (:37, tau_levels :146, rows :203-227)
    "sphere_re266k": ("ball1m", {"basic": {"surface_resolution": 25, "flow": {"velocity": 4.0}}}),
    # the shipped ball1m case = RESULTS_SPHERE_RE10M.txt / CASES/ball1m/RESULTS/*.csv
    "sphere_re10m": ("ball1m", None),
    # config 1: coarsest single-level grid, 500 steps
    "ball1m_coarse": ("ball1m", {"basic": {"surface_resolution": 7, "num_levels": 1, "simulation": {"steps": 500}},
                                 "advanced": {"diagnostics": {"freq": 100}}}),
    "wing5": ("Wing_5_deg", None),
    # config 4 at a size the CPU oracle can afford: same case file (symmetric half model, WMLES, inlet turbulence,
    # 63 196 triangles), 3 levels, 300 steps
    "wing5_small": ("Wing_5_deg", {"basic": {"surface_resolution": 150, "num_levels": 3, "simulation": {"steps": 300, "ramp_steps": 400}},
                                   "advanced": {"diagnostics": {"freq": 100}}}),
    "bunny": ("Stanford_bunny", None),
    # config 5's case file at a size the CPU oracle can afford (3 levels, 2.5 M cells, 25 825 Bouzidi cells, wall model, inlet turbulence)
    "bunny_small": ("Stanford_bunny", {"basic": {"surface_resolution": 80, "num_levels": 3}}),
    # config 5: the bunny scaled to fine resolution: 6 levels, 339 M cells, 9.4 G cell-updates per coarse step
    "bunny_fine": ("Stanford_bunny", {"basic": {"surface_resolution": 1300, "num_levels": 6}}),
}


def case_dir(case: str) -> str:
    for base in SEARCH:
        p = os.path.join(base, case)
        if os.path.isfile(os.path.join(p, "config.yaml")):
            return p
    raise FileNotFoundError(f"case folder {case!r} not found in {SEARCH}")


def have_case(case: str) -> bool:
    try:
        case_dir(case)
        return True
    except FileNotFoundError:
        return False
