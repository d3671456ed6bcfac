"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` export:
stall samples and executed warp instructions per CUDA source line."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(open(path)))
out = []; fpath = ""; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] not in ("",):
        try:
            ln = int(r[0])
        except ValueError:
            continue
        i_s = hdr.index("# Samples"); i_i = hdr.index("Instructions Executed")
        try:
            out.append((fpath, ln, r[1].strip(), int(r[i_s]), int(r[i_i])))
        except ValueError:
            pass
ts = sum(o[3] for o in out); ti = sum(o[4] for o in out)
print(f"total samples {ts}, total warp instructions {ti}")
print("--- by samples")
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{o[0]}:{o[1]:4d} {100*o[3]/ts:5.1f}% smp {100*o[4]/ti:5.1f}% inst  {o[2][:110]}")
print("--- by instructions")
for o in sorted(out, key=lambda o: -o[4])[:top]:
    print(f"{o[0]}:{o[1]:4d} {100*o[3]/ts:5.1f}% smp {100*o[4]/ti:5.1f}% inst  {o[2][:110]}")
