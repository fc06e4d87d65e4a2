import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from vpm_import import load
vpm = load()
from oracle import oracle
from helpers import relerr
ROWS = {"X": slice(0, 3), "Gamma": slice(3, 6), "sigma": slice(6, 7), "U": slice(9, 12), "J": slice(15, 24),
        "M": slice(27, 36), "C": slice(36, 37), "SFS": slice(39, 42)}
for f, g, transposed, sfs in ((0.25, 0.25, False, True), (0.25, 0.25, True, True), (0.0, 0.2, False, True), (0.25, 0.25, False, False)):
    pf = vpm.fields.ring_field(Nphi=60, nc=1, kernel=vpm.gaussianerf)
    pf.particles[42, 5:pf.np:37] = 1.0
    pf.transposed = transposed
    ref = pf.particles.copy(order="F")
    rf = vpm.ResidentField(pf)
    rf.UJ(sfs=sfs, reset=True, reset_sfs=sfs)
    oracle.uj_direct(ref, pf.np, "gaussianerf", sfs=sfs, reset=True, reset_sfs=sfs, transposed=transposed)
    rf.download()
    print(f, g, transposed, sfs, "after UJ:", {k: relerr(pf.particles[r, :pf.np], ref[r, :pf.np]) for k, r in ROWS.items() if k in ("U", "J", "SFS")})
    kw = dict(integration="rungekutta3", f=f, g=g, sfs=sfs, Cs=1.0, clip_backscatter=True, relaxation="pedrizzetti", relax=True)
    rf.upload()
    ref = pf.particles.copy(order="F")
    rf.nextstep(1e-2, **kw)
    oracle.field_step(ref, pf.np, "gaussianerf", 1e-2, transposed=transposed, **kw)
    rf.download()
    print("   after step:", {k: float("%.2e" % relerr(pf.particles[r, :pf.np], ref[r, :pf.np])) for k, r in ROWS.items()})
