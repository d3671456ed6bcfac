"""Drop-in module for ``LinearMPCOverNetworks.TrackingMPC`` of the reference (re-export)."""
from rtmpc_b200.mpc import TrackingMPC  # noqa: F401
