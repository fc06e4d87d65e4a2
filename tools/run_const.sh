python tools/variant_sweep.py 131072 winckelmans 21 sfs 2>&1 | tail -3
python tools/variant_sweep.py 262144 winckelmans 21,12 sfs 2>&1 | tail -6
python - <<'PY'
import sys, time; sys.path.insert(0, ".")
from vpm_import import vpm
import numpy as np
for n, kern in ((4900, vpm.winckelmans), (4900, vpm.gaussianerf), (33800, vpm.gaussianerf)):
    pf = vpm.fields.ring_field(Nphi=100, nc=3, kernel=kern) if n == 4900 else vpm.fields.ring_field(Nphi=100, nc=6, R=0.7906, Rcross=0.07906, rings=2, dZ=0.7906, kernel=kern)
    h = vpm.get_handle()
    h.check(h.lib.vpm_pin_host(h.ptr, pf.particles.ctypes.data, pf.particles.nbytes))
    for sfs in (False, True):
        vpm.UJ_direct(pf, sfs=sfs, reset=True, reset_sfs=sfs)
        t = time.perf_counter()
        for _ in range(20): vpm.UJ_direct(pf, sfs=sfs, reset=True, reset_sfs=sfs)
        dt = (time.perf_counter() - t) / 20
        print(pf.np, kern.name, "sfs" if sfs else "uj ", f"{dt*1e3:.3f} ms/call", {k: round(v, 3) for k, v in h.timing().items() if k.endswith("_ms")})
PY
