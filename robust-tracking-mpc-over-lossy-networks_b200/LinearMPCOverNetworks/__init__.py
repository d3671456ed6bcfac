"""Import-path compatibility: ``import LinearMPCOverNetworks.TubeTrackingMPC as TubeTrackMPC`` etc.
resolve to the B200-native classes of :mod:`rtmpc_b200` (put this directory's parent on sys.path)."""
