"""tools/ncu_hot.py <source-page.csv> [min-samples]: instructions of an ncu source page (ncu -i x.ncu-rep --page source --csv)
with their stall samples and top stall reasons; 100 samples ~ one issue slot per loop iteration when the loop body's
plain instructions show ~100."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 150
ia, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
st = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base = int(data[0][0], 16)
tot = sum(int(r[ia]) for r in data)
print("total samples", tot)
for r in data:
    if int(r[ia]) >= thr:
        reasons = sorted(((int(r[i]), hdr[i][6:]) for i in st if r[i] not in ('', '0')), reverse=True)[:3]
        print("%5x %6d %4.1f%% ex=%7d %-56s %s" % (int(r[0], 16) - base, int(r[ia]), 100 * int(r[ia]) / tot, int(r[ie]), r[1].strip()[:56], reasons))
