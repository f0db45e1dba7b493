"""Top stall-sample SASS lines per kernel from `ncu --page source --csv` output (stdin)."""
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
kern, hdr, rows = None, None, []


def flush():
    if not rows:
        return
    tot = sum(r[1] for r in rows) or 1
    print(f"\n=== {kern[:110]}  (total samples {tot}, {len(rows)} instrs)")
    for i, (src, s, reason, execd) in sorted(enumerate(rows), key=lambda t: -t[1][1])[:top]:
        print(f"  {100 * s / tot:5.1f}%  #{i:5d}  exec={execd:>8}  {src.strip()[:90]:90s} {reason}")


for rec in csv.reader(sys.stdin):
    if not rec:
        continue
    if rec[0] == "Kernel Name":
        flush()
        kern, hdr, rows = rec[1], None, []
        continue
    if rec[0] == "Address":
        hdr = rec
        si = hdr.index("# Samples")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        ei = hdr.index("Instructions Executed")
        continue
    if hdr is None:
        continue
    try:
        s = int(rec[si])
    except Exception:
        continue
    best = max(stall_cols, key=lambda c: int(rec[c[0]] or 0))
    rows.append((rec[1], s, best[1] if int(rec[best[0]] or 0) else "", rec[ei]))
flush()
