"""Drop-in module for ``LinearMPCOverNetworks.TubeRegulatorMPC`` of the reference (re-export)."""
from rtmpc_b200.mpc import TubeRegulatorMPC  # noqa: F401
