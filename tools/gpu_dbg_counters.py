"""Development counters of the active-set solver inside the rollout kernel (library built with -DRTMPC_AS_DEBUG, e.g.
RTMPC_NVCC_EXTRA="-DRTMPC_AS_DEBUG" python -m rtmpc_b200.build): one benchmark rollout, then the counters
(rtmpc_as.cuh: g_as_dbg)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "robust-tracking-mpc-over-lossy-networks_b200"))
import bench                                      # noqa: E402
from rtmpc_b200 import _lib                        # noqa: E402
from rtmpc_b200.rollout import RemoteLoop          # noqa: E402

L = _lib.lib()
if not hasattr(L, "rtmpc_debug_counters"):
    raise SystemExit("library built without -DRTMPC_AS_DEBUG")
mpc, Z = bench.build_controller(extended=False)
B, T = 4096, 250
loop = RemoteLoop(mpc, B, kind="tube", w_half=bench.HW, Z=Z)
p = np.array([0.1 * (i % 10) for i in range(B)])
buf = (C.c_ulonglong * 64)()
loop.reset()
loop.run(T, bench.REF, p_loss=p, seed=679)
L.rtmpc_debug_counters(buf, 1)
loop.reset()
loop.run(T, bench.REF, p_loss=p, seed=680)
import torch
torch.cuda.synchronize()
L.rtmpc_debug_counters(buf, 0)
c = list(buf)
names = {0: "solves past the unconstrained test", 1: "carried", 2: "moved", 3: "changes", 5: "from-scratch inversions",
         6: "their candidates", 7: "GI adds", 8: "GI drops", 9: "warm multiplier drops", 10: "certifications that formed the rows from G' z",
         11: "certifications", 12: "tier 0 tried (accumulated row values)", 13: "tier 1 tried (factored tables)"}
for k, v in names.items():
    print(f"[{k:2d}] {v:55s} {c[k]}")
print("working-set size histogram:", c[16:48])
