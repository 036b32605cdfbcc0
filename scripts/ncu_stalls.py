"""Summarise the per-instruction stall samples of an ncu report (source page, SASS):
python scripts/ncu_stalls.py <report.ncu-rep> [kernel index] [top n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = {"name": line, "rows": []}
        blocks.append(cur)
    elif cur is not None and line.startswith('"'):
        cur["rows"].append(line)
# (the CSV source page lists every captured launch twice - SASS view and source-correlated view carry the same rows -:
# collapse exact consecutive duplicates so that block i is launch i of the capture)
dedup = []
for blk in blocks:
    if dedup and dedup[-1]["name"] == blk["name"] and dedup[-1]["rows"] == blk["rows"]:
        continue
    dedup.append(blk)
blocks = dedup
b = blocks[kidx]
print(b["name"][:160])
rows = list(csv.reader(b["rows"]))
hdr, data = rows[0], rows[1:]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
print("total samples", tot, " instructions", len(data))
agg = {h: sum(int(r[col[h]] or 0) for r in data) for h in stall_cols}
print("stall mix:", ", ".join(f"{h[6:]}={v * 100 // max(1, tot)}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 100 // max(1, tot) >= 1))
order = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = data[i]
    s = int(r[col["# Samples"]] or 0)
    reasons = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {s * 100.0 / max(1, tot):5.1f}%  exec={r[col['Instructions Executed']]:>9}  {r[col['Source']].strip()[:90]:<90} {reasons[0][1]}:{reasons[0][0]} {reasons[1][1]}:{reasons[1][0]}")
