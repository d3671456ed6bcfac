"""Drop-in module for ``LinearMPCOverNetworks.RegulatorMPC`` of the reference (re-export)."""
from rtmpc_b200.mpc import RegulatorMPC  # noqa: F401
