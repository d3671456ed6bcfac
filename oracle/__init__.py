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

PARITY STATUS: **parity unpinned against Clarabel itself.**  The arithmetic of
the reference lives in third-party packages (cvxpy>=1.4.1 -> clarabel, polytope
>=0.2.4, control>=0.9.3.post2, scipy linprog) of which only scipy is present in
this image; the reference ships no tests, golden vectors or fixtures for the QP.
What *does* pin this oracle:
  * Darup-Teichrib ``k_star = 5 / 6 / 10`` for eps = 1e-1 / 1e-2 / 1e-3
    ("Examples of Set Operations/Example of Approximation of mRPI_Darup.py":50-55);
  * the run-time invariants of the example scripts (tube containment, x - x_hat in Z
    when Theta_t = 1, exact estimate for the Pezzutto scheme, u in U);
  * an independent solver cross-check of the QP oracle (scipy SLSQP / trust-constr
    on the reference's un-condensed formulation) -- the QP is strictly convex in
    (x, u, x_bar, u_bar) so its minimiser is unique and any solver converged well
    below Clarabel's tol_gap=1e-7 reproduces Clarabel's answer to that tolerance.
"""
