"""CPU oracle for the tube-MPC hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain numpy/scipy restatement of what the reference
(EricssonResearch/Robust-Tracking-MPC-over-Lossy-Networks) computes on the
path named in BASELINE.json: QP assembly + solve, packet/actuator/estimator
state machines, plant step, and the set computations that produce the QP data.

Rules:
  * Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
    ``--impl reference`` legs may import anything from here.  The product
    package never does; it fails loudly when its CUDA library is missing.
  * Every function cites the reference file:line it follows.

PARITY STATUS: **pinned to the reference's own code for everything but the numerical QP solver.**
``oracle/refshim`` runs the reference's modules and scripts unmodified in the build container (stand-ins for the absent
third-party packages ``polytope``, ``control``, ``cvxpy`` only); ``tests/golden/make_reference_fixtures.py`` writes
``tests/golden/ref_*.npz`` from them; ``tests/test_reference_pin.py`` (CPU) holds the oracle to those fixtures and -
where ``/root/reference`` exists - to the live modules; ``tests/test_gpu_reference.py`` holds the CUDA path to them.
Rows of SURVEY 8(a):
  * A1-A3, E1-E2 (actuators, estimators): reference classes themselves, integers bit-exact, floats to round-off;
  * S1-S5 (support, Pontryagin difference, Rakovic, Darup, terminal set): the reference's ``utils_polytope`` functions
    and the classes' ``setup_optimization`` on the stand-in set algebra; Darup ``k_star = 5 / 6 / 10`` as shipped;
  * Q1, Q3-Q6 (problem statements, packets, class logic incl. the G2 quirk): the reference's
    ``generate_optimization_problem`` / ``determine_packet`` / ``encapsulate`` executed unmodified, the stated problem
    compared entry by entry with the oracle's;
  * whole closed loops of configs 1-3 in the scripts' own call and random-number order.
**Unpinned: Q2, Clarabel's floating-point answer** (cvxpy>=1.4.1 -> clarabel, absent: no wheel, no network).  The
solver under the ``cvxpy`` stand-in is this oracle's interior-point method with a certified polish; the QPs are strictly
convex in (x, u, x_bar, u_bar), so the minimiser Clarabel approximates to tol_gap = 1e-7 is the point the oracle and the
CUDA kernels compute to round-off (every golden solution also carries an independent KKT / NNLS certificate and an SLSQP
cross-check, ``tests/test_oracle_qp.py``).
"""
