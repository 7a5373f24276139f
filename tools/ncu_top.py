"""Summarise the per-instruction page of an ncu report: hottest SASS lines by stall samples.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_top.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
tot_inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
print("total samples", tot, "instructions executed", tot_inst, "SASS lines", len(body))
order = sorted(range(len(body)), key=lambda k: -int(body[k][ix["# Samples"]] or 0))[:top]
for k in sorted(order):
    r = body[k]
    print("%4d %6.2f%% exec=%-10s %s" % (k, 100.0 * int(r[ix["# Samples"]] or 0) / max(tot, 1),
                                        r[ix["Instructions Executed"]], r[ix["Source"]][:100]))
