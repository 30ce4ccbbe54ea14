"""Case configuration and physical -> lattice scaling: host-side mirror of the kept Julia driver files
``config_loader.jl`` (keys + defaults, :119-196) and ``physics_scaling.jl`` (:59-176).

The case format (``CASES/<name>/config.yaml``) is the reference's, verbatim.  All arithmetic is Float64 in the
reference's operation order; values the reference stores as Float32 (u_lattice, tau_min, sponge_thickness, ...)
are rounded through ``np.float32`` at the same places.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Any, Optional

import numpy as np
import yaml


def _get(d: Any, *keys, default=None):
    """safe_get (config_loader.jl:98-107)."""
    cur = d
    for i, k in enumerate(keys):
        if cur is None or not isinstance(cur, dict) or k not in cur:
            if default is not None:
                return default
            raise KeyError("Missing config key: " + " -> ".join(keys[: i + 1]))
        cur = cur[k]
    return default if (cur is None and default is not None) else cur


def _f32(x) -> float:
    return float(np.float32(x))


@dataclass
class CaseConfig:
    case_dir: str
    stl_file: str
    stl_scale: float
    surface_resolution: int
    num_levels_config: int
    symmetric: bool
    reference_area_full_model: float
    reference_area: float
    reference_chord: float
    reference_length_for_meshing: float
    reference_dimension: str
    fluid_density: float
    fluid_kinematic_viscosity: float
    flow_velocity: float
    steps: int
    ramp_steps: int
    output_freq: int
    u_target: float            # Float32 value
    c_wale: float              # Float32
    tau_min: float             # Float32
    inlet_turbulence: float    # Float32
    nu_sgs_background: float   # Float32
    sponge_blend_distributions: bool
    temporal_interpolation: bool
    auto_levels: bool
    max_levels: int
    min_coarse_blocks: int
    wall_model_enabled: bool
    domain_upstream: float
    domain_downstream: float
    domain_lateral: float
    domain_height: float
    sponge_thickness: float    # Float32
    block_size_config: int
    refinement_margin: int
    refinement_strategy: str
    wake_enabled: bool
    wake_length: float
    wake_width_factor: float
    wake_height_factor: float
    boundary_method: str
    bouzidi_levels: int
    q_min_threshold: float     # Float32
    force_enabled: bool
    moment_center: list
    diag_freq: int
    gpu_async_depth: int
    raw: dict = field(default_factory=dict, repr=False)


def _deep_update(dst: dict, src: dict):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _deep_update(dst[k], v)
        else:
            dst[k] = v


def load_case_configuration(case_dir: str, overrides: Optional[dict] = None) -> CaseConfig:
    """config_loader.jl:109-196.  ``overrides`` is a nested dict merged over the YAML (used by the tests to run
    the golden logs' configuration ``surface_resolution: 25, velocity: 14.8`` on the shipped case file)."""
    with open(os.path.join(case_dir, "config.yaml")) as fh:
        cfg = yaml.safe_load(fh)
    if overrides:
        _deep_update(cfg, overrides)
    sym = bool(_get(cfg, "advanced", "refinement", "symmetric_analysis", default=False))
    area_full = float(_get(cfg, "basic", "reference_area_of_full_model", default=0.0))
    return CaseConfig(
        case_dir=case_dir,
        stl_file=_get(cfg, "basic", "stl_file"),
        stl_scale=float(_get(cfg, "basic", "stl_scale")),
        surface_resolution=int(_get(cfg, "basic", "surface_resolution")),
        num_levels_config=int(_get(cfg, "basic", "num_levels")),
        symmetric=sym,
        reference_area_full_model=area_full,
        reference_area=area_full / 2.0 if sym else area_full,
        reference_chord=float(_get(cfg, "basic", "reference_chord", default=0.0)),
        reference_length_for_meshing=float(_get(cfg, "basic", "reference_length_for_meshing", default=0.0)),
        reference_dimension=str(_get(cfg, "basic", "reference_dimension", default="x")),
        fluid_density=float(_get(cfg, "basic", "fluid", "density", default=1.225)),
        fluid_kinematic_viscosity=float(_get(cfg, "basic", "fluid", "kinematic_viscosity", default=1.5e-5)),
        flow_velocity=float(_get(cfg, "basic", "flow", "velocity", default=10.0)),
        steps=int(_get(cfg, "basic", "simulation", "steps")),
        ramp_steps=int(_get(cfg, "basic", "simulation", "ramp_steps")),
        output_freq=int(_get(cfg, "basic", "simulation", "output_freq")),
        u_target=_f32(_get(cfg, "advanced", "numerics", "u_lattice", default=0.01)),
        c_wale=_f32(_get(cfg, "advanced", "numerics", "c_wale", default=0.20)),
        tau_min=_f32(_get(cfg, "advanced", "numerics", "tau_min", default=0.505)),
        inlet_turbulence=_f32(_get(cfg, "advanced", "numerics", "inlet_turbulence_intensity", default=0.01)),
        nu_sgs_background=_f32(_get(cfg, "advanced", "numerics", "nu_sgs_background", default=0.0005)),
        sponge_blend_distributions=bool(_get(cfg, "advanced", "numerics", "sponge_blend_distributions", default=True)),
        temporal_interpolation=bool(_get(cfg, "advanced", "numerics", "temporal_interpolation", default=True)),
        auto_levels=bool(_get(cfg, "advanced", "high_re", "auto_levels", default=False)),
        max_levels=int(_get(cfg, "advanced", "high_re", "max_levels", default=12)),
        min_coarse_blocks=int(_get(cfg, "advanced", "high_re", "min_coarse_blocks", default=4)),
        wall_model_enabled=bool(_get(cfg, "advanced", "high_re", "wall_model", "enabled", default=False)),
        domain_upstream=float(_get(cfg, "advanced", "domain", "upstream", default=0.75)),
        domain_downstream=float(_get(cfg, "advanced", "domain", "downstream", default=1.5)),
        domain_lateral=float(_get(cfg, "advanced", "domain", "lateral", default=0.75)),
        domain_height=float(_get(cfg, "advanced", "domain", "height", default=0.75)),
        sponge_thickness=_f32(_get(cfg, "advanced", "domain", "sponge_thickness", default=0.10)),
        block_size_config=int(_get(cfg, "advanced", "refinement", "block_size", default=8)),
        refinement_margin=int(_get(cfg, "advanced", "refinement", "margin", default=2)),
        refinement_strategy=str(_get(cfg, "advanced", "refinement", "strategy", default="geometry_first")),
        wake_enabled=bool(_get(cfg, "advanced", "refinement", "wake_enabled", default=False)),
        wake_length=float(_get(cfg, "advanced", "refinement", "wake_length", default=0.25)),
        wake_width_factor=float(_get(cfg, "advanced", "refinement", "wake_width_factor", default=0.1)),
        wake_height_factor=float(_get(cfg, "advanced", "refinement", "wake_height_factor", default=0.1)),
        boundary_method=str(_get(cfg, "advanced", "boundary", "method", default="bouzidi")),
        bouzidi_levels=int(_get(cfg, "advanced", "boundary", "bouzidi_levels", default=1)),
        q_min_threshold=_f32(_get(cfg, "advanced", "boundary", "q_min_threshold", default=0.001)),
        force_enabled=bool(_get(cfg, "advanced", "forces", "enabled", default=True)),
        moment_center=list(_get(cfg, "advanced", "forces", "moment_center", default=[0.25, 0.0, 0.0])),
        diag_freq=int(_get(cfg, "advanced", "diagnostics", "freq", default=500)),
        gpu_async_depth=int(_get(cfg, "advanced", "gpu", "async_depth", default=8)),
        raw=cfg,
    )


@dataclass
class DomainParameters:
    """physics_scaling.jl:14-57."""
    num_levels: int
    mesh_min: tuple
    mesh_max: tuple
    mesh_center: tuple
    mesh_extent: tuple
    reference_length: float
    reference_chord: float
    reference_area: float
    moment_center: tuple
    domain_size: tuple
    mesh_offset: tuple
    dx_fine: float
    dx_coarse: float
    dx_levels: list
    nx_coarse: int
    ny_coarse: int
    nz_coarse: int
    bx_max: int
    by_max: int
    bz_max: int
    tau_levels: list          # Float32 values
    re_number: float
    u_physical: float
    rho_physical: float
    nu_physical: float
    length_scale: float
    time_scale: float
    velocity_scale: float
    force_scale: float
    tau_fine: float
    wall_model_active: bool


def compute_domain_from_mesh(cfg: CaseConfig, mesh_min, mesh_max) -> DomainParameters:
    """physics_scaling.jl:86-176."""
    mesh_center = tuple((mesh_min[i] + mesh_max[i]) / 2 for i in range(3))
    mesh_extent = tuple(mesh_max[i] - mesh_min[i] for i in range(3))
    if cfg.reference_length_for_meshing > 0:
        ref_length = cfg.reference_length_for_meshing
    else:
        ref_length = {"x": mesh_extent[0], "y": mesh_extent[1], "z": mesh_extent[2]}.get(cfg.reference_dimension, max(mesh_extent))
    ref_chord = cfg.reference_chord if cfg.reference_chord > 0 else mesh_extent[0]
    if cfg.reference_area > 0:
        ref_area = cfg.reference_area
    else:
        ref_area = mesh_extent[1] * mesh_extent[2] * 2 if cfg.symmetric else mesh_extent[1] * mesh_extent[2]
    mc_rel = tuple(float(v) for v in cfg.moment_center)
    u_phys, nu_phys, rho_phys = cfg.flow_velocity, cfg.fluid_kinematic_viscosity, cfg.fluid_density
    re_number = u_phys * ref_length / nu_phys
    # compute_tau_for_levels (:59-62)
    nu_lattice_fine0 = float(cfg.u_target) * cfg.surface_resolution / re_number
    tau_fine = max(3.0 * nu_lattice_fine0 + 0.5, float(cfg.tau_min))

    domain_x = ref_length * (cfg.domain_upstream + cfg.domain_downstream) + mesh_extent[0]
    domain_y = (mesh_max[1] + ref_length * cfg.domain_lateral) if cfg.symmetric else (mesh_extent[1] + 2 * ref_length * cfg.domain_lateral)
    domain_z = mesh_extent[2] + 2 * ref_length * cfg.domain_height
    dx_fine = ref_length / cfg.surface_resolution
    min_domain = min(domain_x, domain_y, domain_z)
    ratio = min_domain / (dx_fine * cfg.min_coarse_blocks * cfg.block_size_config)   # :71-74
    max_levels_domain = 1 if ratio < 1.0 else int(math.floor(1 + math.log2(ratio)))
    if cfg.num_levels_config > 0:
        num_levels = min(cfg.num_levels_config, max_levels_domain)
    else:
        num_levels = min(max_levels_domain, cfg.max_levels) if cfg.auto_levels else min(8, max_levels_domain)
    dx_coarse = dx_fine * 2 ** (num_levels - 1)
    dx_levels = [dx_fine * 2 ** (num_levels - lvl) for lvl in range(1, num_levels + 1)]
    bs = cfg.block_size_config
    nx_coarse = max(bs, int(math.ceil(math.ceil(domain_x / dx_coarse) / bs) * bs))
    ny_coarse = max(bs, int(math.ceil(math.ceil(domain_y / dx_coarse) / bs) * bs))
    nz_coarse = max(bs, int(math.ceil(math.ceil(domain_z / dx_coarse) / bs) * bs))
    domain_x, domain_y, domain_z = nx_coarse * dx_coarse, ny_coarse * dx_coarse, nz_coarse * dx_coarse
    bx_max, by_max, bz_max = nx_coarse // bs, ny_coarse // bs, nz_coarse // bs
    mesh_x = ref_length * cfg.domain_upstream
    mesh_y = 0.0 if cfg.symmetric else (domain_y / 2 - mesh_center[1])
    mesh_z = domain_z / 2 - mesh_center[2]
    mesh_offset = (mesh_x - mesh_min[0], mesh_y, mesh_z)
    length_scale = dx_fine
    velocity_scale = u_phys / float(cfg.u_target)
    time_scale = length_scale / velocity_scale
    tau_levels = []
    for lvl in range(1, num_levels + 1):
        tau_lvl = tau_fine if lvl == num_levels else 0.5 + (tau_fine - 0.5) * 2.0 ** (num_levels - lvl)
        tau_levels.append(_f32(tau_lvl))
    force_scale = rho_phys * length_scale ** 4 / time_scale ** 2
    moment_center = (mesh_min[0] + mesh_offset[0] + mc_rel[0] * ref_chord,
                     mesh_center[1] + mesh_offset[1] + mc_rel[1] * ref_chord,
                     mesh_center[2] + mesh_offset[2] + mc_rel[2] * ref_chord)
    return DomainParameters(
        num_levels=num_levels, mesh_min=tuple(mesh_min), mesh_max=tuple(mesh_max), mesh_center=mesh_center,
        mesh_extent=mesh_extent, reference_length=ref_length, reference_chord=ref_chord, reference_area=ref_area,
        moment_center=moment_center, domain_size=(domain_x, domain_y, domain_z), mesh_offset=mesh_offset,
        dx_fine=dx_fine, dx_coarse=dx_coarse, dx_levels=dx_levels, nx_coarse=nx_coarse, ny_coarse=ny_coarse,
        nz_coarse=nz_coarse, bx_max=bx_max, by_max=by_max, bz_max=bz_max, tau_levels=tau_levels, re_number=re_number,
        u_physical=u_phys, rho_physical=rho_phys, nu_physical=nu_phys, length_scale=length_scale, time_scale=time_scale,
        velocity_scale=velocity_scale, force_scale=force_scale, tau_fine=tau_fine, wall_model_active=cfg.wall_model_enabled)
