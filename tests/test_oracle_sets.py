"""The oracle's set computations against the reference's own known answers and invariants."""
import os

import numpy as np
import pytest

import helpers as H  # noqa: F401  (path setup)
from oracle import ref_numerics as rn
from oracle import ref_sets as rs
from oracle import ref_setup as su
from oracle.ref_polytope import Polytope, extreme, reduce


def _darup_system():
    # "Examples of Set Operations/Example of Approximation of mRPI_Darup.py":17-47
    A = np.array([[1.0, 1.0], [0.0, 1.0]])
    B = np.array([[0.5], [1.0]])
    W = su.box([0.1, 0.1])
    X = Polytope(np.r_[np.eye(2), -np.eye(2)], np.r_[4.0, 2.0, 8.0, 4.0])
    U = su.box([1.0])
    K, _ = rn.dlqr(A, B, np.eye(2), np.eye(1))
    return A - B @ K, W, X, U, K


@pytest.mark.parametrize("eps,k_expected", [(1e-1, 5), (1e-2, 6), (1e-3, 10)])
def test_darup_k_star_known_answers(eps, k_expected):
    """The only numeric known answers the reference states (same file, :50-55)."""
    Acl, W, X, U, K = _darup_system()
    rpi, C, status, k_star = rs.darup_rpi(Acl, W, X, U, K, eps, 50)
    assert status == 0 and k_star == k_expected
    assert rpi.A.shape[0] == k_expected * 6


def test_darup_rpi_is_invariant():
    Acl, W, X, U, K = _darup_system()
    rpi, _, _, _ = rs.darup_rpi(Acl, W, X, U, K, 1e-2, 50)
    V = extreme(reduce(rpi))
    VW = extreme(W)
    img = (V @ Acl.T)[:, None, :] + VW[None, :, :]
    assert np.all(img.reshape(-1, 2) @ rpi.A.T <= rpi.b + 1e-9)


def test_support_and_pont_diff_on_boxes():
    P1, P2 = su.box([3.0, 2.0]), su.box([0.5, 0.25])
    assert abs(rs.support(P1, np.array([1.0, 1.0])) - 5.0) < 1e-9
    D = rs.pont_diff(P1, P2)
    assert np.allclose(D.b, [2.5, 1.75, 2.5, 1.75])


def test_double_integrator_sets_match_fixture():
    s = H.load("sets_di.npz")
    d = su.tube_tracking_setup(**su.double_integrator(), fixed_initial_state=True)
    assert d["Z"].A.shape == (48, 2) and d["Xf"].A.shape == (26, 5) and d["t_star"] == 8
    assert np.allclose(d["Xc"].b, s["Xc_b"], atol=1e-12) and np.allclose(d["Uc"].b, s["Uc_b"], atol=1e-12)
    assert np.allclose(d["K"], s["K"]) and np.allclose(d["P"], s["P"])
    # Rakovic invariance: A Z + W inside Z (up to the epsilon of the approximation)
    Z, W = d["Z"], su.double_integrator()["W"]
    V = extreme(Z)
    img = (V @ d["Acl"].T)[:, None, :] + extreme(W)[None, :, :]
    assert np.all(img.reshape(-1, 2) @ Z.A.T <= Z.b + 1e-9)


def test_dlyap_convention_G1():
    """P solves  Acl P Acl' - P + Ql = 0  (python-control convention the reference relies on)."""
    c = su.linear_cartpole()
    K, P, Acl = rn.lqr_terminal_data(c["A"], c["B"], c["Q"], c["R"])
    Ql = c["Q"] + K.T @ c["R"] @ K
    assert np.abs(Acl @ P @ Acl.T - P + Ql).max() < 1e-6 * np.abs(P).max()
    assert np.allclose(np.diag(P), [80197.7, 332990.8, 10425.7, 507608.6], rtol=1e-5)


def test_cartpole_fixture_facts():
    s = H.load("sets_cp.npz")
    assert s["Z_A"].shape == (854, 4)
    assert np.allclose(s["Xc_b"][:4], [4.44190075, 3.82958828, 0.07586331, 1.15946501], atol=1e-7)
    assert np.allclose(s["Uc_b"], 3.91068269, atol=1e-7)
    assert s["Xf_A"].shape[1] == 9


@pytest.mark.skipif(not os.environ.get("RTMPC_SLOW"), reason="13 min of LPs; set RTMPC_SLOW=1")
def test_cartpole_sets_regenerate():
    s = H.load("sets_cp.npz")
    d = su.tube_tracking_setup(**su.linear_cartpole(), fixed_initial_state=True)
    assert d["Z"].A.shape == s["Z_A"].shape and d["Xf"].A.shape == s["Xf_A"].shape
    assert np.allclose(d["Xf"].b, s["Xf_b"])
