

python - <<'PY'
import sys, time, os; sys.path.insert(0, ".")
from vpm_import import load; vpm = load()
import numpy as np, torch
from flowvpm_jl_b200 import sharding
h = vpm.get_handle()
n = 262144
pf = vpm.fields.cloud_field(n)
src8 = torch.from_numpy(np.ascontiguousarray(vpm.source_system_to_buffer(pf).T)).cuda()
f = sharding.ShardedField(h, src8, n, 0, 1, vpm.gaussianerf.id)
for flags, name in ((0, "shortcut"), (16, "no-shortcut")):
    f.uj(flags); torch.cuda.synchronize()
    f.uj(flags); torch.cuda.synchronize()
    ms = h.timing()["uj_ms"]
    print("gaussianerf", name, f"{ms:.2f} ms {n*n/ms/1e6:.1f} G/s")
for nn, kern in ((4900, vpm.gaussianerf), (33800, vpm.gaussianerf)):
    pf = vpm.fields.ring_field(Nphi=100, nc=3, kernel=kern) if nn == 4900 else vpm.fields.ring_field(Nphi=100, nc=6, R=0.7906, Rcross=0.07906, rings=2, dZ=0.7906, kernel=kern)
    h.check(h.lib.vpm_pin_host(h.ptr, pf.particles.ctypes.data, pf.particles.nbytes))
    vpm.UJ_direct(pf)
    t = time.perf_counter()
    for _ in range(20): vpm.UJ_direct(pf)
    dt = (time.perf_counter() - t) / 20
    print(pf.np, kern.name, f"{dt*1e3:.3f} ms/call", {k: round(v, 3) for k, v in h.timing().items() if k.endswith("_ms")})
PY
