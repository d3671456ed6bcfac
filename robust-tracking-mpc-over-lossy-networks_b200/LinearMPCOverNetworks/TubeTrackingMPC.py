"""Drop-in module for ``LinearMPCOverNetworks.TubeTrackingMPC`` of the reference (re-export)."""
from rtmpc_b200.mpc import TubeTrackingMPC, ExtendedTubeTrackingMPC  # noqa: F401
