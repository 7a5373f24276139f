"""Decode the scheduling control fields (stall count, write / read scoreboard, wait mask) of an `nvdisasm -hex` listing.

    nvdisasm -hex x.cubin | python tools/sass_ctrl.py [substring of the function name]

sm_70+ encoding: bits 105..125 of the 128-bit instruction = stall[4] yield[1] wbar[3] rbar[3] wait[6] reuse[4].
"""
import re, sys
pat = sys.argv[1] if len(sys.argv) > 1 else None
cur, show, pend = None, pat is None, None
for line in sys.stdin:
    if line.startswith(".text."):
        show = pat is None or pat in line
        if show: print(line.rstrip())
        continue
    if not show: continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", line)
    if m:
        pend = (m.group(1), m.group(2).strip()); continue
    m = re.match(r"\s*/\* (0x[0-9a-f]+) \*/", line)
    if m and pend:
        hi = int(m.group(1), 16)
        c = hi >> 41
        stall, yld, wb, rb, wait = c & 15, (c >> 4) & 1, (c >> 5) & 7, (c >> 8) & 7, (c >> 11) & 63
        print("%s  st%-2d %s W%s R%s wait[%s]  %s" % (pend[0], stall, "Y" if yld else " ", "-" if wb == 7 else wb, "-" if rb == 7 else rb,
              "".join(str(i) for i in range(6) if wait >> i & 1), pend[1]))
        pend = None
    elif line.strip().startswith(".L_"):
        print(line.strip())
