"""Top SASS instructions by warp-stall samples for the N-th kernel of an .ncu-rep (needs --import-source on).
usage: python tools/ncu_hotspots.py file.ncu-rep [kernel_index=0] [top=30]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and row and row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] is not None and len(row) == len(cur["hdr"]):
        cur["rows"].append(row)
b = blocks[kidx]
h = b["hdr"]
si, ai, ni = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(float(r[ai] or 0) for r in b["rows"])
print(b["name"][:120], "total samples", tot, "instructions", len(b["rows"]))
for r in sorted(b["rows"], key=lambda r: -float(r[ai] or 0))[:top]:
    stalls = sorted(((float(r[i] or 0), h[i]) for i in stall_cols), reverse=True)[:2]
    st = ",".join(f"{n[6:]}={v:.0f}" for v, n in stalls if v > 0)
    print(f"{float(r[ai]) / tot * 100:5.1f}%  n={r[ni]:>8s}  {r[si][:90]:90s} {st}")
