"""Shared-memory wavefronts per source line: python scripts/ncu_smem.py file.ncu-rep [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
for row in csv.reader(io.StringIO(raw)):
    if not row: continue
    if row[0] == "File Path": cur = row[1].split("/")[-1]; continue
    if row[0] == "Line No": hdr = row; continue
    if hdr is None or len(row) != len(hdr) or row[0] == "": continue
    d = dict(zip(hdr, row))
    try:
        wf, ideal, inst = int(d["L1 Wavefronts Shared"]), int(d["L1 Wavefronts Shared Ideal"]), int(d["Instructions Executed"])
    except Exception:
        continue
    a = agg[(cur, int(row[0]))]; a[0] += wf; a[1] += ideal; a[2] += inst; a[3] = row[1].strip()[:100]
tot = sum(a[0] for a in agg.values())
print("total shared wavefronts", tot)
for (f, ln), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{100*a[0]/max(tot,1):5.1f}%  wf={a[0]:9d} ideal={a[1]:9d} inst={a[2]:9d}  {f}:{ln:<4d} {a[3]}")
