timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q 2>&1 | tail -3
python - <<'PY'
import sys, time; sys.path.insert(0, ".")
from vpm_import import load; vpm = load()
h = vpm.get_handle()
for n in (262144,):
    pf = vpm.fields.cloud_field(n, kernel=vpm.singular)
    for pin in (False, True):
        if pin: h.check(h.lib.vpm_pin_host(h.ptr, pf.particles.ctypes.data, pf.particles.nbytes))
        for sfs in (False, True):
            vpm.UJ_direct(pf, sfs=sfs, reset=True, reset_sfs=sfs)
            t = time.perf_counter(); vpm.UJ_direct(pf, sfs=sfs, reset=True, reset_sfs=sfs); dt = time.perf_counter() - t
            tm = h.timing()
            print(n, "pinned" if pin else "pageable", "sfs" if sfs else "uj", f"call {dt*1e3:.1f} ms: h2d {tm['h2d_ms']:.2f} uj {tm['uj_ms']:.1f} sfs {tm['sfs_ms']:.1f} d2h {tm['d2h_ms']:.2f}")
PY
