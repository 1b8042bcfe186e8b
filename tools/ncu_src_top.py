"""Summarise an `ncu --page source --csv` export: stall-reason totals and the hottest instructions."""
import csv
import sys

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
sect = int(sys.argv[3]) if len(sys.argv) > 3 else 0   # which kernel section of the export
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = starts[sect]
end = starts[sect + 1] if sect + 1 < len(starts) else len(rows)
hdr = rows[h]
print(rows[h - 1][:2] if h else "")
data = [r for r in rows[h + 1:end] if len(r) == len(hdr) and r[0].startswith("0x")]
iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall = [(n, i) for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[iS]) for r in data) or 1
print("samples", tot, "instructions", len(data))
agg = sorted(((sum(int(r[i] or 0) for r in data), n) for n, i in stall), reverse=True)
print("  ".join(f"{n[6:]}={100 * v / tot:.1f}%" for v, n in agg[:9]))
for r in sorted(data, key=lambda r: -int(r[iS]))[:topn]:
    st = sorted(((int(r[i] or 0), n[6:]) for n, i in stall), reverse=True)[:2]
    print(f"{100 * int(r[iS]) / tot:5.1f}% ex={r[iEx]:>9} {r[iSrc].strip()[:64]:64s} {st}")
