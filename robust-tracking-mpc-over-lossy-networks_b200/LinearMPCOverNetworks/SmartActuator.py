"""Drop-in module for ``LinearMPCOverNetworks.SmartActuator`` of the reference (re-export)."""
from rtmpc_b200.local_remote import SmartActuator, ConsistentActuator  # noqa: F401
