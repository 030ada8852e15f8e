"""Read a PRD_FLASH_TRACE dump (prd_triattn.cu) and print per-phase durations and the timeline of one CTA.
Usage: python tools/flash_trace.py trace.bin [t0 t1]"""
import sys
import collections
import numpy as np

NEV, NCTA, NW = 512, 8, 12
EV = {1: "start", 10: "c:or", 11: "c:S", 12: "c:pr0", 13: "c:pr1", 14: "c:commit", 20: "item", 21: "s_ready",
      22: "passA_end", 23: "O_read", 24: "token", 25: "passB_end", 30: "lastO", 31: "stored", 40: "ch0", 41: "ch1", 42: "ch2", 43: "ch3", 44: "ch4", 45: "ch5", 46: "ch6", 47: "ch7"}
a = np.fromfile(sys.argv[1], dtype=np.int64).reshape(NCTA, NW, NEV)


def events(c, w):
    r = a[c, w]
    r = r[r != 0]
    return [(int(x >> 8), int(x & 255)) for x in r]


def stats(warps, title):
    dur = collections.defaultdict(list)
    for c in range(NCTA):
        for w in warps:
            ev = events(c, w)
            for (t0, e0), (t1, e1) in zip(ev, ev[1:]):
                dur[(EV.get(e0, e0), EV.get(e1, e1))].append(t1 - t0)
    print(title, "(cycles): mean / p10 / p90 / count")
    for k, v in sorted(dur.items(), key=lambda x: -np.sum(x[1])):
        v = np.array(v)
        print(f"  {k[0]:>10s} -> {k[1]:10s} {v.mean():8.0f} {np.percentile(v,10):8.0f} {np.percentile(v,90):8.0f} {len(v):6d}")


stats(range(0, 8), "softmax warps")
stats((8, 9), "UMMA warps")
t_lo = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
t_hi = int(sys.argv[3]) if len(sys.argv) > 3 else 42000
c = 1
base = events(c, 0)[0][0]
merged = []
cols = {0: 0, 4: 1, 8: 2, 9: 3}
for w, col in cols.items():
    merged += [(t - base, col, f"w{w}:{EV.get(e, e)}") for t, e in events(c, w)]
merged.sort()
print(f"\nCTA {c} timeline: group A warp 0 | group B warp 4 | UMMA A | UMMA B")
for t, col, s in merged:
    if t_lo <= t <= t_hi:
        print(f"{t:8d} " + " " * (18 * col) + s)
