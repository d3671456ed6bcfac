"""SASS-level view of an `ncu --page source --csv --print-source cuda,sass` export of the rollout kernel:
hot code footprint (bytes of SASS that carry 50/90/99 % of the executed warp instructions), stall totals by reason,
and the instructions on which instruction-fetch stalls (stall_no_inst) concentrate."""
import csv, sys, collections
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = None
ins = {}      # address -> dict
cur_line = None; cur_file = ""
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None: continue
    if r[0].isdigit(): cur_line = (cur_file, int(r[0])); continue
    if len(r) < len(hdr) or not r[2].startswith("0x"): continue
    a = int(r[2], 16)
    def g(name):
        try: return int(r[ix[name]])
        except Exception: return 0
    d = ins.setdefault(a, dict(sass=r[3].strip(), line=cur_line, ex=0, smp=0, st=collections.Counter()))
    d["ex"] += g("Instructions Executed"); d["smp"] += g("# Samples")
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h: d["st"][h] += g(h)
addrs = sorted(ins)
base = addrs[0]
tot_ex = sum(d["ex"] for d in ins.values()); tot_smp = sum(d["smp"] for d in ins.values())
print("SASS instructions:", len(addrs), "bytes:", (addrs[-1] - base + 16), "executed warp instr:", tot_ex, "samples:", tot_smp)
by = sorted(ins.values(), key=lambda d: -d["ex"])
c = 0; k = 0; marks = [0.5, 0.9, 0.99]
for d in by:
    c += d["ex"]; k += 1
    while marks and c >= marks[0] * tot_ex:
        print(f"  {marks[0]*100:.0f} % of executed instructions in {k} instructions = {k*16/1024:.1f} KB"); marks.pop(0)
st = collections.Counter()
for d in ins.values(): st.update(d["st"])
print("stall samples by reason:", ", ".join(f"{k[6:]} {v/tot_smp*100:.1f}%" for k, v in st.most_common(8)))
# 128-byte lines touched, weighted
lines = collections.Counter()
for a in addrs: lines[(a - base) // 128] += ins[a]["ex"]
hot_lines = [l for l, v in lines.items() if v > 0.0002 * tot_ex]
print("128-B instruction lines with > 0.02 % of the executed instructions each:", len(hot_lines), "=", len(hot_lines) * 128 / 1024, "KB")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print("top stall_no_inst instructions:")
for a in sorted(addrs, key=lambda a: -ins[a]["st"]["stall_no_inst"])[:n]:
    d = ins[a]
    print(f"  +{a-base:#07x} no_inst {d['st']['stall_no_inst']:6d} smp {d['smp']:6d} ex {d['ex']:10d}  {d['line']}  {d['sass'][:60]}")
if len(sys.argv) > 3:
    # static size and dynamic weight per source function-ish bucket (file, 25-line bucket)
    b = collections.defaultdict(lambda: [0, 0, 0, 0])
    for a in addrs:
        d = ins[a]
        f, l = d["line"] if d["line"] else ("?", 0)
        k = (f, l // 25 * 25)
        b[k][0] += 1; b[k][1] += d["ex"]; b[k][2] += d["st"]["stall_no_inst"]; b[k][3] += 1 if d["ex"] > 1e-5 * tot_ex else 0
    print("file:line-bucket  static instrs (hot)  executed %  no_inst %")
    tn = sum(v[2] for v in b.values())
    for k, v in sorted(b.items(), key=lambda kv: -kv[1][1]):
        if v[1] > 0.003 * tot_ex or v[0] > 100:
            print(f"  {k[0]}:{k[1]:4d}  {v[0]:5d} ({v[3]:5d})  {v[1]/tot_ex*100:5.1f}  {v[2]/tn*100:5.1f}")
