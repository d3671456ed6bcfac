"""Drop-in module for ``LinearMPCOverNetworks.Estimator`` of the reference (re-export)."""
from rtmpc_b200.local_remote import Estimator, RobustEstimator  # noqa: F401
