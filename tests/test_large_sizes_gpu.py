"""Parity at BASELINE sizes.  The CPU oracle needs ~8 s per step at 256^3 on 16 threads, so:
  * 256^3 (16.8 M cells): two full oracle steps vs CUDA strict, bit-exact, and vs CUDA fast within tolerance;
  * 512^3 (134 M cells, the bench configuration): size-independent properties —
      - strict and fast agree within FP32 round-off after 6 steps (fields downloaded and compared cell by cell),
      - flow statistics are consistent (n_fluid = all cells, min <= mean <= max),
      - block-order invariance: uploading the same state through a PERMUTED reference block order gives the same
        per-cell result (checks the reference-layout <-> internal-layout maps at full size).
"""
import numpy as np
import pytest

from open_ludwig_b200 import cabi
from open_ludwig_b200.host import synthetic as syn
from util import default_params, load_state, rel_err_rho_u

pytestmark = pytest.mark.gpu


def _run(lib, lv, state, p, steps):
    with cabi.Context(lib) as c:
        c.add_level(lv)
        load_state(c, 0, *state)
        c.step_batch(1, steps, 0.03, p)
        c.sync()
        return {"rho": c.download(0, cabi.RHO), "vel": c.download(0, cabi.VEL if steps % 2 else cabi.VEL_TEMP)}, c.flow_stats(0)


def test_256_cube_oracle_parity(oracle_lib, cuda_lib):
    nb = 32
    lv = syn.make_box_level(nb, nb, nb)
    state = syn.noise_state(lv)
    cells = (nb * 8,) * 3
    ref, _ = _run(oracle_lib, lv, state, default_params(cells, strict=1), 2)
    got, _ = _run(cuda_lib, lv, state, default_params(cells, strict=1), 2)
    for k in ref:
        assert np.array_equal(ref[k].view(np.int32), got[k].view(np.int32)), k
    fast, _ = _run(cuda_lib, lv, state, default_params(cells, strict=0), 2)
    e_rho, e_u = rel_err_rho_u(ref, fast)
    assert e_rho <= 1e-5 and e_u <= 1e-5, (e_rho, e_u)


def test_512_cube_properties(cuda_lib):
    nb = 64
    lv = syn.make_box_level(nb, nb, nb)
    state = syn.noise_state(lv)
    cells = (nb * 8,) * 3
    strict, s1 = _run(cuda_lib, lv, state, default_params(cells, strict=1), 6)
    fast, s0 = _run(cuda_lib, lv, state, default_params(cells, strict=0), 6)
    e_rho, e_u = rel_err_rho_u(strict, fast)
    assert e_rho <= 1e-5 and e_u <= 1e-5, (e_rho, e_u)
    for s in (s0, s1):
        assert s["n_fluid"] == lv.n_cells and s["rho_min"] <= s["rho_mean"] <= s["rho_max"]
    assert abs(s0["rho_mean"] - s1["rho_mean"]) < 1e-7 and abs(s0["kinetic_energy"] / s1["kinetic_energy"] - 1) < 1e-5
    del strict
    # block-order invariance: present the same level with its blocks listed in a random reference order
    rng = np.random.default_rng(5)
    perm = rng.permutation(lv.n_blocks)                       # new reference index i holds old block perm[i]
    inv = np.empty_like(perm); inv[perm] = np.arange(lv.n_blocks)
    lv2 = syn.make_box_level(nb, nb, nb)
    lv2.active_block_coords = lv.active_block_coords[perm]
    nt = lv.neighbor_table[:, perm]
    lv2.neighbor_table = np.where(nt > 0, inv[np.maximum(nt, 1) - 1] + 1, 0).astype(np.int32)
    bp = lv.block_pointer
    lv2.block_pointer = np.where(bp > 0, inv[np.maximum(bp, 1) - 1] + 1, 0).astype(np.int32)
    f, rho, vel = state
    state2 = (f[:, perm], rho[perm], vel[:, perm])
    del f, rho, vel, state
    fast2, _ = _run(cuda_lib, lv2, state2, default_params(cells, strict=0), 6)
    assert np.array_equal(fast2["rho"].view(np.int32), fast["rho"][perm].view(np.int32))
    assert np.array_equal(fast2["vel"].view(np.int32), fast["vel"][:, perm].view(np.int32))
