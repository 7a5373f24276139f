"""Decode the scheduling control fields of a cuobjdump -sass listing (Volta+ encoding) and
print, per instruction, the static stall count / barriers; sums over an address range.
usage: python tools/sass_stalls.py file.sass [start_hex end_hex]"""
import re, sys
lines = open(sys.argv[1]).read().split("\n")
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 60
ins = []
i = 0
while i < len(lines):
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r"\s*/\* (0x[0-9a-f]+) \*/", lines[i + 1])
        if m2:
            ins.append((int(m.group(1), 16), m.group(2).strip(), int(m2.group(1), 16)))
            i += 2
            continue
    i += 1
tot = 0
n = 0
for addr, txt, hiw in ins:
    if not (lo <= addr < hi):
        continue
    stall = (hiw >> 41) & 0xF
    yld = (hiw >> 45) & 1
    wbar = (hiw >> 46) & 7
    rbar = (hiw >> 49) & 7
    wait = (hiw >> 52) & 0x3F
    tot += max(stall, 1)
    n += 1
    if "-v" in sys.argv:
        print("%05x st=%2d y=%d w=%s r=%s wait=%02x  %s" % (addr, stall, yld, wbar if wbar != 7 else "-", rbar if rbar != 7 else "-", wait, txt))
print("instructions %d, sum of static stalls %d cycles" % (n, tot))
