"""Compile csrc/*.cu into librtmpc_b200.so (in-tree, next to this file) with nvcc for sm_100a."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(_HERE), "csrc")
OUT = os.path.join(_HERE, "librtmpc_b200.so")
SOURCES = ["rtmpc_capi.cu", "rtmpc_as.cu", "rtmpc_ipm.cu", "rtmpc_rollout.cu", "rtmpc_lp.cu"]
HEADERS = ["rtmpc_common.cuh", "rtmpc_ipm.cuh", "rtmpc_as.cuh", "rtmpc_loop.cuh", "rtmpc_launch.h",
           os.path.join("..", "..", "include", "rtmpc.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]
OBJ_DIR = os.path.join(os.path.dirname(_HERE), "build")


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """One object per translation unit (compiled in parallel), then one shared library."""
    if not force and not _stale():
        return OUT
    from concurrent.futures import ThreadPoolExecutor
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(OBJ_DIR, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
        # RTMPC_NVCC_EXTRA: extra flags for development builds (e.g. "-DRTMPC_RO_MAXW5=20"); never set by the product
        extra = os.environ.get("RTMPC_NVCC_EXTRA", "").split()
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
        return obj, res.stderr

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        done = list(ex.map(compile_one, SOURCES))
    if verbose:
        for _, log in done:
            print(log)
    res = subprocess.run([nvcc, "-shared", "-o", OUT] + [o for o, _ in done], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
