"""Print the hot SASS lines (and stall summary) of a kernel from `ncu --page source --csv --print-source sass` output."""
import csv, sys, collections
path = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
rows = list(csv.reader(open(path)))
# several kernels may be concatenated: split at header rows
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "hdr": None, "rows": []}; blocks.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = r; continue
    cur["rows"].append(r)
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
b = blocks[which]; hdr = b["hdr"]
print("kernel:", b["name"], "(%d kernels in file)" % len(blocks))
ia = hdr.index("Address"); isrc = hdr.index("Source"); ie = hdr.index("Instructions Executed"); ip = hdr.index("Avg. Predicated-On Threads Executed"); iss = hdr.index("# Samples")
stall = [k for k in range(len(hdr)) if hdr[k].startswith("stall") and "Not" not in hdr[k]]
data = []
for r in b["rows"]:
    try: e = float(r[ie])
    except Exception: continue
    data.append((r, e))
tot = sum(e for _, e in data); mx = max(e for _, e in data)
agg = collections.Counter()
for r, e in data:
    for k in stall:
        try: agg[hdr[k]] += float(r[k])
        except Exception: pass
print("total warp instructions %.4g" % tot, "max line count %.4g" % mx)
print("stalls:", [(k, int(v)) for k, v in agg.most_common(8)])
for r, e in data:
    if e >= thr * mx:
        top = sorted(((float(r[k] or 0), hdr[k][6:]) for k in stall), reverse=True)[0]
        print("%s %8.0f %5s %6s %-12s %s" % (r[ia][-5:], e / 1000, r[ip], r[iss], top[1] if top[0] > 0 else "", r[isrc]))
