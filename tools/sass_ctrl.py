"""development aid: print SASS with the scheduling control fields (stall count, yield, barriers) decoded
from cuobjdump's hex words.  usage: sass_ctrl.py <obj> <kernel-substring> [start_addr end_addr]"""
import re, subprocess, sys
obj, kern = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.splitlines()
on = False; pend = None; total = 0
for ln in txt:
    if "Function :" in ln:
        on = kern in ln
        continue
    if not on: continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", ln)
    if m:
        pend = (int(m.group(1), 16), m.group(2).strip()); continue
    m = re.match(r"\s*/\* (0x[0-9a-f]+) \*/", ln)
    if m and pend:
        w = int(m.group(1), 16)
        stall = (w >> 41) & 0xf; yld = (w >> 45) & 1; wr = (w >> 46) & 7; rd = (w >> 49) & 7; wait = (w >> 52) & 0x3f
        a, ins = pend; pend = None
        if lo <= a <= hi:
            total += stall
            print("%05x  st=%2d %s wr=%s rd=%s wait=%02x  %s" % (a, stall, "Y" if yld else " ", wr if wr != 7 else "-", rd if rd != 7 else "-", wait, ins))
print("sum of stall counts in range:", total)
