"""The C ABI from a plain C host program (examples/c_abi_demo.c): it compiles against include/rtmpc.h with gcc, fails
loudly without a GPU, and on a GPU its packets equal the ones the Python mirror gets for the same instances."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import helpers as H

PKG = os.path.join(H.ROOT, "robust-tracking-mpc-over-lossy-networks_b200", "rtmpc_b200")


def _build(tmp_path):
    from rtmpc_b200 import build
    build.build()
    exe = str(tmp_path / "c_abi_demo")
    blob = str(tmp_path / "qp_cp.bin")
    subprocess.run(["gcc", "-O2", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(H.ROOT, "include"),
                    os.path.join(H.ROOT, "examples", "c_abi_demo.c"), "-o", exe, "-L", PKG, "-lrtmpc_b200",
                    "-Wl,-rpath," + PKG], check=True)
    subprocess.run([sys.executable, os.path.join(H.ROOT, "examples", "dump_qp_desc.py"), blob], check=True,
                   stdout=subprocess.DEVNULL)
    return exe, blob


def test_c_host_program_compiles_and_refuses_to_run_without_a_gpu(tmp_path):
    import torch
    exe, blob = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    res = subprocess.run([exe, blob, "8"], capture_output=True, text=True)
    assert res.returncode == 1 and "no CUDA device" in res.stderr          # no CPU fallback behind the ABI


@pytest.mark.gpu
def test_c_host_program_matches_python_mirror(tmp_path):
    from rtmpc_b200.qp import BatchedQP
    exe, blob = _build(tmp_path)
    B = 512
    res = subprocess.run([exe, blob, str(B)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert re.search(r"batch 512 optimal 512", res.stdout)
    rows = re.findall(r"instance (\d+) x1 (\S+) status (\d+) steps (\d+) u0 (\S+) u_ss (\S+)", res.stdout)
    assert len(rows) >= 4
    s = H.load("sets_cp.npz")
    qp = BatchedQP(H.spec_tube_tracking(s), Kss=s["K"])
    x = np.zeros((B, 4))
    x[:, 0] = 0.45 * np.arange(B) / (B - 1)
    r = np.zeros((B, 4))
    r[:, 0] = 0.5
    _, U, st, _ = qp.solve_host(x, r)
    for b, x1, status, _, u0, uss in rows:
        b = int(b)
        assert float(x1) == x[b, 0] and int(status) == st[b]
        assert float(u0) == U[b, 0, 0] and float(uss) == U[b, -1, 0]       # same library, same bits
