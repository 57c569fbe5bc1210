"""Top stall lines of a kernel from `ncu --page source --csv --print-source sass` output: python tools/top_stalls.py file.csv [kernel#] [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "hdr": None, "rows": []}; blocks.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = r; continue
    cur["rows"].append(r)
b = blocks[which]; hdr = b["hdr"]
ia = hdr.index("Address"); isrc = hdr.index("Source"); ie = hdr.index("Instructions Executed"); iss = hdr.index("# Samples")
stall = [k for k in range(len(hdr)) if hdr[k].startswith("stall")]
data = []
for i, r in enumerate(b["rows"]):
    try:
        top = max(((float(r[k] or 0), hdr[k][6:]) for k in stall))
        data.append((int(r[iss] or 0), i, r[ia][-5:], float(r[ie]), top[1], r[isrc]))
    except Exception:
        pass
print(b["name"], "samples", sum(d[0] for d in data), "warp inst %.4g" % sum(d[3] for d in data))
for d in sorted(data, reverse=True)[:n]:
    print(d)
