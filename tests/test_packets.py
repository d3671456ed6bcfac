"""Wire formats: the reference's packet dicts <-> struct-of-arrays, and foreign H-representations."""
import numpy as np
import pytest

import helpers as H  # noqa: F401  (path setup)
from rtmpc_b200 import _lib, packets as pk
from rtmpc_b200.polytope import Polytope


def test_controller_packets_round_trip_reference_shapes():
    """TubeTrackingMPC.encapsulate (:211-227) / ExtendedTubeTrackingMPC.determine_packet (:351-369) shapes."""
    rng = np.random.default_rng(0)
    nu, N, nx = 2, 5, 3
    pkts = [{"U_t": rng.normal(size=(nu, N + 1)), "q_t": 3, "x_nom_0": rng.normal(size=nx)},
            {"U_t": None, "q_t": 4, "x_nom_0": None},                              # infeasible problem: U_t = None
            {"U_t": rng.normal(size=(nu, N + 1)), "q_t": 0, "x_nom_0": rng.normal(size=nx)}]
    b = pk.pack_controller_packets(pkts)
    assert b["U_t"].shape == (3, N + 1, nu) and b["q_t"].dtype == np.int32 and b["x_nom_0"].shape == (3, nx)
    assert list(b["status"]) == [_lib.OPTIMAL, _lib.INFEASIBLE, _lib.OPTIMAL] and np.isnan(b["U_t"][1]).all()
    assert np.array_equal(b["U_t"][0, :, 1], pkts[0]["U_t"][1])                   # time-major payload
    back = pk.unpack_controller_packets(b)
    for p, q in zip(pkts, back):
        assert q["q_t"] == p["q_t"] and set(q) == set(p)
        if p["U_t"] is None:
            assert q["U_t"] is None and q["x_nom_0"] is None
        else:
            assert q["U_t"].shape == (nu, N + 1) and np.array_equal(q["U_t"], p["U_t"]) and np.array_equal(q["x_nom_0"], p["x_nom_0"])
    # plain (non-extended) packets carry no x_nom_0 key at all
    plain = pk.unpack_controller_packets(pk.pack_controller_packets([{"U_t": np.ones((1, 4)), "q_t": 7}]))
    assert set(plain[0]) == {"U_t", "q_t"} and plain[0]["q_t"] == 7
    with pytest.raises(ValueError):
        pk.pack_controller_packets([{"U_t": None, "q_t": 0}])
    with pytest.raises(ValueError):
        pk.pack_controller_packets([{"U_t": np.ones((1, 4)), "q_t": 0}, {"U_t": np.ones((1, 5)), "q_t": 0}])


def test_plant_packets_round_trip_reference_shapes():
    """SmartActuator.encapsulate (:115-123) and ConsistentActuator.encapsulate_extended (:224-231): column vectors."""
    rng = np.random.default_rng(1)
    pkts = [{"x_t": rng.normal(size=(4, 1)), "s_t": k, "x_nom_t": rng.normal(size=(4, 1))} for k in range(5)]
    b = pk.pack_plant_packets(pkts)
    assert b["x_t"].shape == (5, 4) and b["s_t"].dtype == np.int32 and b["x_nom_t"].shape == (5, 4)
    back = pk.unpack_plant_packets(b)
    for p, q in zip(pkts, back):
        assert q["x_t"].shape == (4, 1) and np.array_equal(q["x_t"], p["x_t"]) and q["s_t"] == p["s_t"]
        assert np.array_equal(q["x_nom_t"], p["x_nom_t"])
    assert set(pk.unpack_plant_packets(pk.pack_plant_packets([{"x_t": np.zeros(2), "s_t": 1}]))[0]) == {"x_t", "s_t"}


class ForeignPolytope:
    """What a third-party H-representation object may look like: un-normalised rows, column ``b``, no other API."""

    def __init__(self, A, b):
        self.A = np.asarray(A, float)
        self.b = np.asarray(b, float).reshape(-1, 1)


def test_foreign_h_representation_is_normalised_like_polytope_package():
    F = ForeignPolytope([[2.0, 0.0], [0.0, -4.0], [0.0, 0.0], [-3.0, 4.0]], [8.0, 2.0, 1.0, 10.0])
    P = pk.as_polytope(F)
    assert isinstance(P, Polytope) and P.A.shape == (3, 2)                        # the zero row is dropped (upstream behaviour)
    assert np.allclose(np.linalg.norm(P.A, axis=1), 1.0) and np.allclose(P.b, [4.0, 0.5, 2.0])
    assert pk.as_polytope(P) is P
    A, b = pk.to_hrep(F)
    assert b.shape == (4,) and A.shape == (4, 2)
    with pytest.raises(TypeError):
        pk.as_polytope(np.eye(2))
    with pytest.raises(ValueError):
        pk.as_polytope(ForeignPolytope(np.eye(2), [1.0, 2.0, 3.0]))


def test_controller_accepts_foreign_constraint_objects_cpu_side():
    """set_state_constraints / set_input_constraints with foreign objects, set pipeline on the injected LP backend (the
    CUDA sweep is covered by tests/test_gpu_reference.py): same tightened sets as with this package's own boxes."""
    from oracle import ref_sets as rs
    from oracle.ref_polytope import Polytope as OP
    from rtmpc_b200 import mpc, sets as up
    up.set_support_backend(lambda V, dirs: np.max(dirs @ V.T, axis=1))
    try:
        A, B = np.array([[1.0, 1.0], [0.0, 1.0]]), np.array([[0.0], [1.0]])
        c = mpc.TubeTrackingMPC(A, B, np.eye(2), np.eye(1), 10)
        c.set_input_constraints(ForeignPolytope([[3.0], [-0.5]], [3.0, 0.5]))                       # |u| <= 1
        c.set_state_constraints(ForeignPolytope(np.r_[2 * np.eye(2), -0.25 * np.eye(2)], [16.0, 16.0, 2.0, 2.0]))  # |x| <= 8
        c.determine_mRPI(ForeignPolytope(np.r_[10 * np.eye(2), -10 * np.eye(2)], np.ones(4)))       # |w| <= 0.1
        c.tighten_constraints()
        s = H.load("sets_di.npz")
        assert c._Z.A.shape == s["Z_A"].shape
        assert np.allclose(c._Xc.b, s["Xc_b"], atol=1e-9) and np.allclose(c._Uc.b, s["Uc_b"], atol=1e-9)
        assert np.allclose(c._Xc.A, s["Xc_A"]) and np.allclose(c._Uc.A, s["Uc_A"])
    finally:
        up.set_support_backend(None)
