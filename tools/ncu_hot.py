"""Hottest SASS instructions (by warp-stall samples) of each kernel in an `ncu --page source --csv` export.
usage: ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv; python tools/ncu_hot.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 28
kern, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1][:70], "hdr": None, "rows": []}
        kern.append(cur)
        continue
    if cur is None:
        continue
    if cur["hdr"] is None:
        cur["hdr"] = r
        continue
    cur["rows"].append(r)
for k in kern:
    h = k["hdr"]
    i_s, i_n, i_e = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    i_l, i_w, i_sh = h.index("stall_long_sb"), h.index("stall_wait"), h.index("stall_short_sb")
    tot = sum(int(r[i_n]) for r in k["rows"])
    print("=====", k["name"], "total samples", tot, "ninstr", len(k["rows"]))
    top = sorted(range(len(k["rows"])), key=lambda j: -int(k["rows"][j][i_n]))[:N]
    for j in sorted(top):
        r = k["rows"][j]
        print(f"{j:5d} {int(r[i_n]):6d} {100*int(r[i_n])/max(tot,1):5.1f}% ex={int(r[i_e]):8d} long={r[i_l]:>5s} "
              f"short={r[i_sh]:>5s} wait={r[i_w]:>5s}  {r[i_s].strip()[:90]}")
