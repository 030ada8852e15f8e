"""Per-kernel census of the Blackwell-native SASS mnemonics in libprd_sm100.so (the evidence B200_PROFILING.md names):
UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor copies), UBLKCP (bulk copies),
SYNCS (mbarrier), MUFU, HMMA (legacy mma.sync: expected 0).  Usage: python tools/sass_census.py > profiles/rNN_sass_census.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "protein_redesign_b200", "libprd_sm100.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "MUFU", "HMMA", "FFMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = cur.replace("(anonymous namespace)::", "").replace("prd::", "")
            cur = re.sub(r"\(.*", "", cur)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for k in KEYS:
                if op.startswith(k):
                    counts[cur][k] += 1
    print("# SASS census of libprd_sm100.so (cuobjdump -sass, sm_100a): instruction counts per kernel\n")
    print("UTCHMMA = tcgen05.mma (kind::f16 / kind::tf32), LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG = cp.async.bulk.tensor (TMA),")
    print("UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, HMMA = warp-level mma.sync (only the head-dim-16 attention of the backward pass: prd_bwd_attn.cu, HMMA.1688.F32.TF32),\n"
          "UTMASTG = cp.async.bulk.tensor store (the GEMM's fp32 epilogue).\n")
    print("| kernel | " + " | ".join(KEYS) + " |")
    print("|---|" + "---:|" * len(KEYS))
    tot = collections.Counter()
    for name, c in sorted(counts.items(), key=lambda kv: -(kv[1]["UTCHMMA"] * 1000 + kv[1]["UTMALDG"] + kv[1]["FFMA"] * 1e-3)):
        tot.update(c)
        print(f"| `{name[:70]}` | " + " | ".join(str(c[k]) for k in KEYS) + " |")
    print("| **total** | " + " | ".join(str(tot[k]) for k in KEYS) + " |")
    print(f"\n{len(counts)} kernels; {sum(1 for c in counts.values() if c['UTCHMMA'])} issue tcgen05.mma, "
          f"{sum(1 for c in counts.values() if c['UTMALDG'])} use TMA tensor loads, {sum(1 for c in counts.values() if c['HMMA'])} use mma.sync.")


if __name__ == "__main__":
    main()
