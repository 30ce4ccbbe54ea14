"""include/ludwig_b200.h consumed from plain C: tests/c/abi_smoke.c is compiled as C99 (-pedantic) and linked against the product
library (no GPU needed for that), and run on the GPU box, where it must drive create / level_create / step / stats / destroy."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "c", "abi_smoke")


def build():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return EXE


def test_header_compiles_as_c99_and_links(cuda_lib):
    exe = build()
    assert os.access(exe, os.X_OK)
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libludwig_b200.so" in out and "not found" not in out.split("libludwig_b200.so")[1].splitlines()[0]


def test_no_gpu_is_a_clean_error_not_a_fallback(cuda_lib):
    """Without a CUDA device the C program gets an error code from ludwig_ctx_create (exit 2) — never a CPU fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    r = subprocess.run([build()], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "ludwig_ctx_create failed" in r.stderr


@pytest.mark.gpu
def test_c_program_runs_the_hot_path(cuda_lib):
    r = subprocess.run([build()], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "backend cuda-sm100a" in r.stdout and "C ABI smoke ok" in r.stdout
