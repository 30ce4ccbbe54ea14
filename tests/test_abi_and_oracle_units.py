"""CPU-only checks: the C-ABI library loads and exports every symbol include/ludwig_b200.h declares (no compute
calls without a GPU), error behaviour of the ABI (through the oracle, which implements the same contract), and
known-answer tests of the oracle's device functions against independent NumPy restatements."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from open_ludwig_b200 import cabi
from open_ludwig_b200.host import synthetic as syn
from util import default_params, fetch_state, load_state

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "ludwig_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ludwig_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_functions() == sorted(cabi.EXPORTED_SYMBOLS)


@pytest.mark.parametrize("which", ["cuda", "oracle"])
def test_library_exports_every_symbol(which, cuda_lib, oracle_lib):
    lib = C.CDLL(cuda_lib if which == "cuda" else oracle_lib)
    for name in header_functions():
        assert hasattr(lib, name), f"{name} not exported"
    lib.ludwig_backend_name.restype = C.c_char_p
    assert lib.ludwig_backend_name() == (b"cuda-sm100a" if which == "cuda" else b"cpu-oracle")


def test_no_cpu_fallback_when_extension_missing(tmp_path):
    with pytest.raises(cabi.LudwigError, match="no CPU fallback"):
        cabi.load_library(str(tmp_path / "libludwig_b200.so"))


def test_product_package_never_references_the_oracle():
    pkg = os.path.join(ROOT, "open_ludwig_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".jl")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "libludwig_oracle" not in text and "oracle/_build" not in text, os.path.join(dirpath, f)


def test_abi_error_behaviour(oracle_lib):
    """Non-zero return + message instead of an exception across the ABI; levels must be created in order."""
    lv = syn.make_box_level(2, 2, 2)
    with cabi.Context(oracle_lib) as c:
        lv.level_id = 2
        with pytest.raises(cabi.LudwigError, match="order"):
            c.add_level(lv)
        lv.level_id = 1
        c.add_level(lv)
        with pytest.raises(cabi.LudwigError):
            c.download(0, cabi.F_OLD)              # no temporal storage on this level
        with pytest.raises(cabi.LudwigError):
            c.flow_stats(3)                        # no such level


def test_lattice_tables():
    """physics_v2.jl:99-117: ordering, weights, opp = 28-k (1-based)."""
    k = np.arange(27)
    assert np.array_equal(syn.CX, k % 3 - 1) and np.array_equal(syn.CZ, k // 9 - 1)
    assert abs(float(syn.W.astype(np.float64).sum()) - 1.0) < 1e-7
    for i in range(27):
        j = 26 - i
        assert (syn.CX[j], syn.CY[j], syn.CZ[j]) == (-syn.CX[i], -syn.CY[i], -syn.CZ[i])


def test_hash_known_answers():
    """murmur3 fmix32 (physics_utils.jl:17-22): published test values of the finaliser."""
    assert int(syn.gpu_hash(np.array([0], np.uint32))[0]) == 0
    assert int(syn.gpu_hash(np.array([1], np.uint32))[0]) == 0x514E28B7
    assert int(syn.gpu_hash(np.array([0xFFFFFFFF], np.uint32))[0]) == 0x81F16F39
    n = syn.gradient_noise(np.arange(1, 50), 7, 3, 1234)
    assert n.dtype == np.float32 and n.min() >= -1.0 and n.max() < 1.0


def test_oracle_step_matches_numpy_restatement(oracle_lib):
    """One K1 step of the oracle on a fully periodic box vs an independent vectorised NumPy restatement of
    physics_kernels.jl:62-354 (pull, moments, WALE from the previous velocities, regularized collision)."""
    f32 = np.float32
    dims = (2, 2, 2)
    coords, bp, nt = syn.box_topology(*dims, periodic_x=True)
    lv = syn.make_box_level(*dims)
    lv.neighbor_table = nt
    f, rho, vel = syn.noise_state(lv)
    p = default_params((16, 16, 16), strict=1)
    with cabi.Context(oracle_lib) as c:
        c.add_level(lv)
        load_state(c, 0, f, rho, vel)
        c.step_batch(2, 1, 0.03, p)          # even step: reads f / vel, writes f_temp / vel_temp
        got = fetch_state(c, 0)
    # --- NumPy restatement on the dense 16^3 grid
    n = 16
    def dense(a):   # [nb,8,8,8] block array (bx major, bz fastest) -> [z,y,x]
        return a.reshape(2, 2, 2, 8, 8, 8).transpose(2, 3, 1, 4, 0, 5).reshape(n, n, n)
    fd = np.stack([dense(f[k]) for k in range(27)])
    ud = [dense(vel[i]) for i in range(3)]
    pulled = np.stack([np.roll(fd[k], (syn.CZ[k], syn.CY[k], syn.CX[k]), (0, 1, 2)) for k in range(27)])
    r = np.zeros((n, n, n), f32); j = [np.zeros((n, n, n), f32) for _ in range(3)]
    for k in range(27):
        r = r + pulled[k]
        j[0] = j[0] + pulled[k] * f32(syn.CX[k]); j[1] = j[1] + pulled[k] * f32(syn.CY[k]); j[2] = j[2] + pulled[k] * f32(syn.CZ[k])
    r = np.maximum(r, f32(0.01)); inv = f32(1) / r
    u = [j[i] * inv for i in range(3)]
    g = [[f32(0.5) * (np.roll(ud[i], -1, 2 - a) - np.roll(ud[i], 1, 2 - a)) for a in range(3)] for i in range(3)]
    gsq = [[g[i][0] * g[0][jj] + g[i][1] * g[1][jj] + g[i][2] * g[2][jj] for jj in range(3)] for i in range(3)]
    tr = (gsq[0][0] + gsq[1][1] + gsq[2][2]) / f32(3)
    Sd = [[(gsq[i][i] - tr) if i == jj else f32(0.5) * (gsq[i][jj] + gsq[jj][i]) for jj in range(3)] for i in range(3)]
    S = [[f32(0.5) * (g[i][jj] + g[jj][i]) for jj in range(3)] for i in range(3)]
    OP1 = Sd[0][0] ** 2 + Sd[1][1] ** 2 + Sd[2][2] ** 2 + f32(2) * (Sd[0][1] ** 2 + Sd[0][2] ** 2 + Sd[1][2] ** 2)
    OP2 = g[0][0] ** 2 + g[1][1] ** 2 + g[2][2] ** 2 + f32(2) * (S[0][1] ** 2 + S[0][2] ** 2 + S[1][2] ** 2)
    den = OP2 * OP2 * np.sqrt(np.maximum(OP2, f32(1e-12))) + OP1 * np.sqrt(np.sqrt(np.maximum(OP1, f32(1e-12))))
    with np.errstate(divide="ignore", invalid="ignore"):
        nu = np.where((OP1 > f32(1e-12)) & (den > f32(1e-12)), (f32(0.5) * f32(0.5)) * (OP1 * np.sqrt(OP1)) / den, f32(0))
    nu = np.maximum(nu, f32(0.0005)).astype(f32)
    omega = f32(1) / np.maximum(f32(lv.tau) + nu * f32(3), f32(0.500001))
    usq = u[0] * u[0] + u[1] * u[1] + u[2] * u[2]
    feq = []
    for k in range(27):
        cu = f32(syn.CX[k]) * u[0] + f32(syn.CY[k]) * u[1] + f32(syn.CZ[k]) * u[2]
        feq.append(r * syn.W[k] * (f32(1) + f32(3) * cu + f32(4.5) * cu * cu - f32(1.5) * usq))
    Pi = {}
    for a, b_, key in ((0, 0, "xx"), (1, 1, "yy"), (2, 2, "zz"), (0, 1, "xy"), (1, 2, "yz"), (2, 0, "zx")):
        c3 = (syn.CX, syn.CY, syn.CZ)
        acc = np.zeros((n, n, n), f32)
        for k in range(27):
            acc = acc + (pulled[k] - feq[k]) * f32(c3[a][k]) * f32(c3[b_][k])
        Pi[key] = acc
    cs2 = f32(1) / f32(3)
    out = []
    for k in range(27):
        cx, cy, cz = f32(syn.CX[k]), f32(syn.CY[k]), f32(syn.CZ[k])
        reg = syn.W[k] * f32(4.5) * (Pi["xx"] * (cx * cx - cs2) + Pi["yy"] * (cy * cy - cs2) + Pi["zz"] * (cz * cz - cs2)
                                     + f32(2) * (Pi["xy"] * cx * cy + Pi["yz"] * cy * cz + Pi["zx"] * cz * cx))
        out.append(feq[k] + (f32(1) - omega) * reg)
    want = np.stack(out)
    got_d = np.stack([dense(got["f_temp"][k]) for k in range(27)])
    assert np.array_equal(want.view(np.int32), got_d.view(np.int32))          # same FP32 operation order -> same bits
    assert np.array_equal(dense(got["rho"]).view(np.int32), r.view(np.int32))


def _stats_with_nan(lib, where):
    """flow statistics of a 2^3-block box at rest with one NaN planted in rho (where = 'rho') or vel (where = 'vel')"""
    import numpy as np
    from open_ludwig_b200 import cabi
    from open_ludwig_b200.host import synthetic as syn
    lv = syn.make_box_level(2, 2, 2)
    f, rho, vel = syn.noise_state(lv)
    if where == "rho":
        rho[3, 1, 2, 5] = np.nan
    elif where == "vel":
        vel[1, 6, 7, 0, 3] = np.nan
    with cabi.Context(lib) as c:
        c.add_level(lv)
        c.upload(0, cabi.RHO, rho); c.upload(0, cabi.VEL, vel)
        return c.flow_stats(0)


@pytest.mark.parametrize("where", ["none", "rho", "vel"])
def test_oracle_flow_stats_propagate_nan(oracle_lib, where):
    """Julia's minimum / maximum propagate NaN (diagnostics.jl:70-77): rho_min is the console table's only divergence indicator."""
    import math
    s = _stats_with_nan(oracle_lib, where)
    assert s["n_fluid"] == 8 * 512
    assert math.isnan(s["rho_min"]) == (where == "rho") and math.isnan(s["rho_max"]) == (where == "rho")
    assert math.isnan(s["v_max"]) == (where == "vel")
    assert math.isnan(s["rho_mean"]) == (where == "rho") and math.isnan(s["kinetic_energy"]) == (where != "none")


@pytest.mark.gpu
@pytest.mark.parametrize("where", ["none", "rho", "vel"])
def test_cuda_flow_stats_propagate_nan(cuda_lib, oracle_lib, where):
    import math
    s, r = _stats_with_nan(cuda_lib, where), _stats_with_nan(oracle_lib, where)
    for k in s:
        assert (math.isnan(s[k]) and math.isnan(r[k])) or s[k] == pytest.approx(r[k], rel=1e-12), (k, s[k], r[k])
