"""Instruction mix and stall hot spots of one kernel from `ncu -i X.ncu-rep --page source --csv > X.csv`.
python profiles/summarize_ncu_source.py X.csv [kernel-substring]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and len(r) > 5:
        cur["rows"].append(r)
for s in secs:
    if want not in s["name"]:
        continue
    h = s["hdr"]
    ie, src, ss = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    tot = sum(int(r[ie]) for r in s["rows"])
    smp = sum(int(r[ss]) for r in s["rows"])
    print(f"## {s['name']}\nwarp instructions executed: {tot}, stall samples: {smp}\n")
    mix, st = Counter(), Counter()
    for r in s["rows"]:
        t = r[src].split()
        op = t[0] if not t[0].startswith("@") else t[1]
        key = ".".join(op.split(".")[:2]) if op.startswith(("MUFU", "F2FP", "SYNCS", "LDTM", "UTC")) else op.split(".")[0]
        mix[key] += int(r[ie])
        st[key] += int(r[ss])
    print("| opcode | executed | share | stall samples |\n|---|---|---|---|")
    for k, v in mix.most_common(16):
        print(f"| {k} | {v} | {100.0 * v / tot:.1f} % | {st[k]} |")
    print("\nhottest instructions by stall samples:\n")
    top = sorted(s["rows"], key=lambda r: -int(r[ss]))[:12]
    for r in top:
        print(f"    {int(r[ss]):6d} samples  {int(r[ie]):9d} exec   {r[src].strip()[:100]}")
    print()
