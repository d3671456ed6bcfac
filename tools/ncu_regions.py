"""Region shares (stall samples / executed warp instructions) of the rollout kernel from an
`ncu --page source --csv --print-source cuda,sass` export; regions are line ranges of the CUDA sources, found by marker
strings so that they follow the code as it moves."""
import csv, re, sys, os
path = sys.argv[1]
SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "robust-tracking-mpc-over-lossy-networks_b200", "csrc")


def line_of(fname, marker, nth=1):
    k = 0
    for i, l in enumerate(open(os.path.join(SRC, fname)), 1):
        if marker in l:
            k += 1
            if k == nth:
                return i
    raise SystemExit(f"marker not found: {fname}: {marker}")


A = "rtmpc_as.cuh"
marks = [("reductions / keys", A, "__device__ __forceinline__ unsigned long long as_key"),
         ("mat-vec", A, "double as_matvec(int Mo"),
         ("bordering", A, "void as_border(int Mo"),
         ("down-date", A, "void as_downdate(int Mo"),
         ("Gauss-Jordan inversion", A, "unsigned as_invert(int Mo"),
         ("marks / misc", A, "__device__ __forceinline__ void as_mark"),
         ("GI: search + step logic", A, "__device__ __forceinline__ int as_gi("),
         ("GI: row streaming", A, "// row values move by  c W[p][:]"),
         ("GI: border / drop bookkeeping", A, "if (apply_only) { apply_only = false; break; }"),
         ("certification: refinement", A, "__device__ __forceinline__ int as_certify("),
         ("certification: exact rows + checks", A, "// exact row values at z:"),
         ("set-up (parameters, z_u, rows at z_u)", A, "__device__ __forceinline__ int as_solve_instance("),
         ("warm start / refactorisation", A, "// M starts empty"),
         ("GI / certification driver", A, "// ---- 2./3. Goldfarb-Idnani"),
         ("outputs (packet payload, z, warm record)", A, "// ---- outputs ----"),
         ("end", A, "__device__ __forceinline__ int as_pack_iters")]
bounds = [(name, f, line_of(f, m)) for name, f, m in marks]
rows = list(csv.reader(open(path)))
per = {}
fpath, hdr = "", None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit():
        try:
            smp, ins = int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        ln = int(r[0])
        if fpath == A:
            name = "as.cuh header (ld2, accessors)"
            for (n0, _, l0), (_, _, l1) in zip(bounds[:-1], bounds[1:]):
                if l0 <= ln < l1:
                    name = n0
        elif fpath == "rtmpc_loop.cuh":
            name = "closed-loop step + tube check (rtmpc_loop.cuh)"
        elif fpath == "rtmpc_rollout.cu":
            name = "rollout driver (rtmpc_rollout.cu)"
        else:
            name = "intrinsics (sync, REDUX, shuffles, rcp)"
        a = per.setdefault(name, [0, 0])
        a[0] += smp; a[1] += ins
ts = sum(v[0] for v in per.values()); ti = sum(v[1] for v in per.values())
print(f"total stall samples {ts}, warp instructions {ti}")
print("| region | warp instructions | stall samples |\n|---|---|---|")
for k, v in sorted(per.items(), key=lambda kv: -kv[1][0]):
    print(f"| {k} | {100*v[1]/ti:.1f} % | {100*v[0]/ts:.1f} % |")
