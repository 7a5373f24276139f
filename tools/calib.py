"""Step-cost / inter-band lag calibration of the single-pair kernels (DESIGN.md 5 model):
T = tau * (m + 63 + (nbands-1) * lag).  One band gives tau, a few bands give the lag.
usage: python tools/calib.py [rows]"""
import ctypes as C
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ["TSTA_PSA_ALLOW_TALL"] = "1"
from tsta_b200 import api, synth

m = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
CLK = 1.965e6   # cycles per ms
y = synth.random_read(m, 2)
L = api.load()
for tr, cs in ((False, (3, 4)), (True, (4,))):
    for ch in ("32", "16"):
        os.environ["TSTA_PSA_CH"] = ch
        for Cc in cs:
            os.environ["TSTA_PSA_C16"] = str(Cc)
            rows = []
            for nb in (1, 2, 11, 101, 401):
                n = 64 * Cc * nb
                x = synth.random_read(n, 1)
                p = api.PsaPlan(tr)
                prm = api.Params(2, -5, -2, -4)
                api._check(L.tsta_psa_plan_upload(p._h, C.byref(prm), x, n, y, m))
                p.n, p.m = n, m
                for _ in range(2):
                    p.run()
                f = min(p.run()[0] for _ in range(4))
                info = p.info()
                rows.append((info["bands"], f * CLK))
                print("trace=%d CH=%s C=%d bands=%d n=%d m=%d: fill %.3f ms = %.1f cycles/row"
                      % (tr, ch, Cc, info["bands"], n, m, f, f * CLK / m), flush=True)
                p.close()
            tau = rows[0][1] / (m + 63)
            for nb, cyc in rows[1:]:
                print("   tau %.1f cycles/step (1 band); %d bands: lag %.1f steps per band"
                      % (tau, nb, (cyc / tau - m - 63) / (nb - 1)), flush=True)
