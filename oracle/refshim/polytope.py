"""Oracle (test infrastructure): stand-in for the absent third-party package ``polytope`` (pin ``polytope>=0.2.4``,
reference ``setup.cfg:16``) -- exactly the names the reference touches, all backed by ``oracle/ref_polytope.py``:

* ``pc.Polytope(A, b[, vertices=])``, ``.A``, ``.b``, ``.vertices``, ``.copy()``, ``.intersect()``, ``==``, ``in``
  (``utils_polytope.py:37,48,86,134,175,239,259-261``; ``TubeRegulatorMPC.py:101``; ``TubeTrackingMPC.py:57``)
* ``pc.extreme`` (``utils_polytope.py:48,58,145,241``), ``pc.qhull`` (``:167``), ``pc.reduce`` (``TubeRegulatorMPC.py:74``)
* ``pc.polytope`` -- the upstream sub-module, used only inside annotations (``utils_polytope.py:270``)
"""
import sys

from oracle.ref_polytope import (ABS_TOL, Polytope, bounding_box, cheby_ball, extreme, is_fulldim, is_subset,  # noqa: F401
                                 qhull, reduce)

polytope = sys.modules[__name__]


def _plot(self, *args, **kwargs):
    """upstream ``Polytope.plot`` draws with matplotlib; the example scripts call it, nothing reads the result."""
    return None


Polytope.plot = _plot


def box2poly(box):
    """upstream ``pc.box2poly([[lo, hi], ...])`` ("Examples of Set Operations/Example of Several Set Operations.py":13)."""
    import numpy as np
    box = np.asarray(box, dtype=float)
    n = box.shape[0]
    return Polytope(np.r_[np.eye(n), -np.eye(n)], np.r_[box[:, 1], -box[:, 0]])
