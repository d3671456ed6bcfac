"""rtmpc_b200 -- B200-native batched remote tube-MPC (hot path of LinearMPCOverNetworks).

Host-side mirror of the reference's class surface on top of ``librtmpc_b200.so`` (hand-written
sm_100a CUDA behind the C ABI in ``include/rtmpc.h``).  Importing the package does not need a GPU;
every compute entry point does, and raises instead of falling back to the CPU.
"""
__version__ = "0.1.0"
