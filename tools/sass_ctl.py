"""Decode per-instruction scheduling control (stall count, yield, scoreboards) from `cuobjdump -sass` output.
Usage: cuobjdump -sass x.o | python tools/sass_ctl.py [start_regex] [count]"""
import re
import sys

lines = sys.stdin.read().split("\n")
out = []
i = 0
while i < len(lines):
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r"\s*/\* (0x[0-9a-f]+) \*/", lines[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            out.append((m.group(2).strip(), (hi >> 41) & 0xF, (hi >> 45) & 1, (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 0x3F))
            i += 2
            continue
    i += 1
pat = sys.argv[1] if len(sys.argv) > 1 else None
cnt = int(sys.argv[2]) if len(sys.argv) > 2 else 80
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
start = 0
if pat:
    idx = [k for k, o in enumerate(out) if re.search(pat, o[0])]
    start = idx[min(skip, len(idx) - 1)] if idx else 0
tot = 0
for o in out[start:start + cnt]:
    tot += o[1]
    print(f"{o[0][:64]:64s} stall={o[1]:2d} y={o[2]} wb={o[3]} rb={o[4]} wait={o[5]:06b}")
print("sum of stall counts:", tot)
