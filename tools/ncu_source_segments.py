"""Summarise `ncu --page source --csv` of one kernel: instructions executed and stall samples between marker instructions
(loads, tensor-core / TMEM ops, barriers, branches).  usage: ncu_source_segments.py file.csv [section]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        sections.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
sec = sections[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
hdr, data = sec["rows"][0], [r for r in sec["rows"][1:] if len(r) > 10]
iS, iI, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
print(sec["name"], "total inst", sum(int(r[iI]) for r in data), "rows", len(data), "samples", sum(int(r[iN]) for r in data))
MARK = ("UTCHMMA", "LDTM", "STTM", "SYNCS", "BAR.", "LDG", "EXIT", "BRA", "UTCBAR", "STG", "LDS", "STS", "HMMA")
last = 0
for k, r in enumerate(data):
    if any(m in r[iS] for m in MARK):
        seg = data[last:k + 1]
        top = sorted(((sum(int(x[i] or 0) for x in seg), hdr[i]) for i in stalls), reverse=True)[:2]
        print(f"{k:5d} inst={sum(int(x[iI]) for x in seg):10d} samp={sum(int(x[iN]) for x in seg):6d} {top[0][1]}={top[0][0]} {top[1][1]}={top[1][0]} | {r[iS].strip()[:56]} x{r[iI]}")
        last = k + 1
