"""Hot spots of an ncu source-page CSV (SASS view): lines sorted by stall samples + totals per opcode.
Usage: ncu -i rep --page source --csv --print-source sass > x.csv ; python tools/ncu_hot.py x.csv [top]"""
import csv
import collections
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot_s = sum(int(r[ix["# Samples"]]) for r in data)
tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data)
print("total samples", tot_s, "total warp inst", tot_i)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
by_op = collections.defaultdict(lambda: [0, 0])
for n, r in enumerate(data):
    op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[ix["Source"]].split()[1]
    by_op[op][0] += int(r[ix["Instructions Executed"]])
    by_op[op][1] += int(r[ix["# Samples"]])
print("-- per opcode: inst%, samples%")
for op, (i, s) in sorted(by_op.items(), key=lambda x: -x[1][0])[:25]:
    print(f"  {op:40s} {100*i/tot_i:6.2f}% {100*s/tot_s:6.2f}%")
print("-- hottest lines")
order = sorted(range(len(data)), key=lambda n: -int(data[n][ix["# Samples"]]))[:top]
for n in sorted(order):
    r = data[n]
    st = {h[6:]: int(r[ix[h]]) for h in stall_cols if int(r[ix[h]]) > 0}
    st = dict(sorted(st.items(), key=lambda x: -x[1])[:3])
    print(f"  {n:5d} {r[ix['Source']].strip()[:60]:60s} smp {int(r[ix['# Samples']]):6d} ({100*int(r[ix['# Samples']])/tot_s:4.1f}%) inst {r[ix['Instructions Executed']]:>9s} {st}")
