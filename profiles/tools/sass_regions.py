"""ncu source-page CSV (ncu -i X.ncu-rep --page source --csv) -> cumulative stall samples between marker instructions
(barriers, cp.async, atomics, global loads/stores): which phase of a long persistent kernel the warps spend their time in."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
isrc, iss = hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[iss]) for r in data if r[iss].isdigit())
print("total samples", tot)
marks, cum = [], 0
for i, r in enumerate(data):
    cum += int(r[iss]) if r[iss].isdigit() else 0
    if re.search(r"BAR\.SYNC|LDGSTS|LDGDEPBAR|DEPBAR|MEMBAR|ATOM|RED\.|CCTL|MUFU\.TANH|MUFU\.EX2|SHFL|STG|LDG|STS", r[isrc]):
        marks.append((i, r[isrc].strip()[:60], cum))
out = []
for i, src, c in marks:
    key = src.split()[0] if not src.startswith('@') else src.split()[1]
    if out and out[-1][1] == key and i - out[-1][2] < 40:
        out[-1][2], out[-1][3], out[-1][4] = i, c, out[-1][4] + 1
    else:
        out.append([i, key, i, c, 1])
pc = 0
for a, key, b, c, n in out:
    print(f"{a:5d}-{b:5d} {key:26s} x{n:3d} cum={100 * c / tot:5.1f}% (+{100 * (c - pc) / tot:4.1f})")
    pc = c
