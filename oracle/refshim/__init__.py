"""Oracle (test infrastructure): run the reference's OWN modules in this image.

``/root/reference/src/LinearMPCOverNetworks`` is pure Python, but four of its imports are third-party packages that
are absent here (no wheels, no network): ``polytope``, ``control``, ``cvxpy`` (-> ``clarabel``).  ``SmartActuator.py``
and ``Estimator.py`` need none of them and import as they are.  For the rest this package provides *stand-ins for the
absent third-party packages only* -- never for reference code:

* ``polytope``  -> ``oracle/refshim/polytope.py``  (names the reference uses: ``Polytope``, ``extreme``, ``qhull``,
                    ``reduce``, the ``polytope.polytope`` sub-module attribute)  backed by ``oracle/ref_polytope.py``
* ``control``   -> ``oracle/refshim/control.py``   (``dlqr``, ``dlyap``, ``ss``, ``c2d``) backed by scipy
* ``cvxpy``     -> ``oracle/refshim/cvxpy.py``     (``Variable``, ``Parameter``, ``quad_form``, ``Minimize``,
                    ``Problem``, ``CLARABEL``): an affine-expression algebra that turns the problem the reference's
                    ``generate_optimization_problem`` *states* into ``min 1/2 z'Pz + q'z, Ez = e, Gz <= h`` and solves it
                    with ``oracle.ref_qp.solve_qp``.  The problem statement is then the reference's, executed
                    unmodified; only the numerical solver is ours (Clarabel is absent -> "solver unpinned").

``reference_modules()`` puts the stand-ins into ``sys.modules`` for the duration of the import, imports the
reference's modules from ``/root/reference/src`` (read-only, never copied) and removes the stand-ins again.
Used by ``tests/golden/make_reference_fixtures.py`` (which writes the committed ``tests/golden/ref_*.npz``) and by the
live CPU test ``tests/test_reference_pin.py`` (skipped when ``/root/reference`` is absent, e.g. on the GPU box).
"""
import contextlib
import importlib
import os
import sys

REFERENCE_SRC = os.environ.get("RTMPC_REFERENCE_SRC", "/root/reference/src")
_SHIMS = ("polytope", "control", "cvxpy")
_REF_MODULES = ("SmartActuator", "Estimator", "utils_polytope", "RegulatorMPC", "TubeRegulatorMPC", "TrackingMPC",
                "TubeTrackingMPC")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_SRC, "LinearMPCOverNetworks", "SmartActuator.py"))


@contextlib.contextmanager
def _shims_installed():
    saved = {k: sys.modules.get(k) for k in _SHIMS}
    try:
        for k in _SHIMS:
            sys.modules[k] = importlib.import_module("oracle.refshim." + k)
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


_cache = {}


def _is_ref_pkg(k):
    return k == "LinearMPCOverNetworks" or k.startswith("LinearMPCOverNetworks.")


@contextlib.contextmanager
def reference_environment(stub_matplotlib=False):
    """Everything an unmodified reference file needs to import here: ``REFERENCE_SRC`` first on ``sys.path`` (our compat
    package has the same import name - the reference's directory must win and no earlier import of the compat package
    may be picked up), the third-party stand-ins in ``sys.modules`` and, for the example / result scripts, an inert
    ``matplotlib`` (``MagicMock``: the scripts only draw with it).  All of it is undone on exit."""
    if not reference_available():
        raise FileNotFoundError(f"reference sources not found under {REFERENCE_SRC}")
    stale = {k: sys.modules.pop(k) for k in list(sys.modules) if _is_ref_pkg(k)}
    sys.modules.update(_cache_modules)
    saved_mpl = {k: sys.modules.get(k) for k in ("matplotlib", "matplotlib.pyplot")}
    sys.path.insert(0, REFERENCE_SRC)
    import warnings
    try:
        with _shims_installed(), warnings.catch_warnings():
            warnings.simplefilter("ignore", SyntaxWarning)      # LaTeX backslashes in the reference's docstrings
            if stub_matplotlib:
                from unittest import mock
                mpl = mock.MagicMock(name="matplotlib")
                mpl.rcParams = {}

                def subplots(nrows=1, ncols=1, **_):
                    n = int(nrows) * int(ncols)
                    axes = mock.MagicMock() if n == 1 else tuple(mock.MagicMock() for _ in range(n))
                    return mock.MagicMock(), axes
                mpl.pyplot.subplots = subplots
                sys.modules["matplotlib"] = mpl
                sys.modules["matplotlib.pyplot"] = mpl.pyplot
            yield
    finally:
        sys.path.remove(REFERENCE_SRC)
        for k in [k for k in sys.modules if _is_ref_pkg(k)]:
            _cache_modules[k] = sys.modules.pop(k)
        sys.modules.update(stale)
        if stub_matplotlib:
            for k, v in saved_mpl.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v


_cache_modules = {}


def reference_modules():
    """dict name -> module of the reference's ``LinearMPCOverNetworks`` package, imported from ``REFERENCE_SRC`` with
    the stand-ins for its absent third-party imports.  Raises ``FileNotFoundError`` when the reference is not there."""
    if _cache:
        return _cache
    with reference_environment():
        for name in _REF_MODULES:
            m = importlib.import_module("LinearMPCOverNetworks." + name)
            assert os.path.realpath(m.__file__).startswith(os.path.realpath(REFERENCE_SRC)), m.__file__
            _cache[name] = m
    return _cache


def run_reference_script(relpath, quiet=True):
    """Execute one of the reference's example / result scripts AS SHIPPED (``runpy.run_path`` on the file where it lies
    under the reference root, nothing patched) and return its module globals.  ``relpath`` is relative to the
    reference root (the parent of ``REFERENCE_SRC``)."""
    import io
    import runpy
    import warnings
    path = os.path.join(os.path.dirname(REFERENCE_SRC), relpath)
    with reference_environment(stub_matplotlib=True), warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)
        if quiet:
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                g = runpy.run_path(path, run_name="__main__")
            g["__stdout__"] = buf.getvalue()
        else:
            g = runpy.run_path(path, run_name="__main__")
    return g
