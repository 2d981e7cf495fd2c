"""Per-source-line stall samples of an ncu capture (needs -lineinfo + --import-source on):
   python scripts/ncu_source.py file.ncu-rep [top_n]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, "", collections.Counter()])
done_kernel = 0
for row in csv.reader(io.StringIO(raw)):
    if not row:
        continue
    if row[0] == "File Path":
        cur_file = row[1].split("/")[-1]; continue
    if row[0] == "Function Name":
        continue
    if row[0] == "Line No":
        hdr = row; continue
    if hdr is None or len(row) != len(hdr) or row[0] == "":
        continue
    d = dict(zip(hdr, row))
    try:
        samples = int(d["# Samples"]); inst = int(d["Instructions Executed"])
    except Exception:
        continue
    key = (cur_file, int(row[0]))
    a = agg[key]
    a[0] += samples; a[1] += inst; a[2] = row[1].strip()[:110]
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k:
            try:
                a[3][k] += int(v)
            except Exception:
                pass
tot = sum(a[0] for a in agg.values())
print("total samples", tot)
for (f, ln), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    st = ", ".join(f"{k[6:]}={v}" for k, v in a[3].most_common(3))
    print(f"{100*a[0]/max(tot,1):5.1f}%  inst={a[1]:9d}  {f}:{ln:<4d} {a[2]}   [{st}]")
