"""Per-source-line dynamic instruction counts with opcode mix from an
`ncu --page source --csv --print-source cuda,sass` export: python tools/ncu_opmix.py file.csv units [top]"""
import csv, re, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
N = float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
hdr = None; cur = None; f = ""
per = collections.OrderedDict(); txt = {}; smp = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": f = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[0] != "":
        try:
            cur = (f, int(r[0])); txt[cur] = r[1].strip(); smp[cur] += int(r[hdr.index("# Samples")])
        except ValueError:
            cur = None
    elif cur and len(r) > 7 and r[2].startswith("0x"):
        try: n = int(r[hdr.index("Instructions Executed")])
        except ValueError: continue
        per.setdefault(cur, []).append((r[3].strip(), n))
tot = sum(n for v in per.values() for _, n in v); ts = sum(smp.values())
print(f"instructions per unit {tot / N:.0f}")
for k, v in sorted(per.items(), key=lambda kv: -sum(n for _, n in kv[1]))[:top]:
    s = sum(n for _, n in v)
    ops = collections.Counter()
    for ins, n in v:
        ins = re.sub(r"^@!?U?P\d+\s+", "", ins); ops[ins.split()[0].split(".")[0]] += n
    print(f"{k[0][6:]}:{k[1]} {s / N:6.0f}/unit {100 * smp[k] / ts:4.1f}%t  {txt[k][:84]}")
    print("      " + ", ".join(f"{o}:{c / N:.0f}" for o, c in ops.most_common(8)))
