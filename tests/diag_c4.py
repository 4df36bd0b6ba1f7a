"""TEST INFRASTRUCTURE (a script, not collected by pytest; it is the only place outside test_*.py / bench.py that touches the oracle).
Diagnostic: C4 sample (200x200 QuadP PhaseField(NeoHooke), plane-strain extension) on the GPU against the oracle: error
distribution relative to the natural scale of each entry, reproducibility, and independence of the K1 launch shape"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle.check import entry_scales  # noqa: E402

dims = tuple(int(x) for x in sys.argv[1:3]) if len(sys.argv) > 2 else (200, 200)
w = bench.Workload("C4", dims, jitter=0.1, device_ctor=(os.environ.get("DIAG_DEVICE_CTOR", "1") == "1"))
ref = w.oracle(nthreads=bench.host_threads())[0]
got = w.product()
got2 = w.product()
for ph, ((phi, r, Kt), (phi0, r0, (cp0, rv0, nz0)), (_, _, Kt2)) in enumerate(zip(got, ref, got2)):
    nz = Kt.nzval
    print("phase", ph, "pattern equal", np.array_equal(Kt.colptr, cp0) and np.array_equal(Kt.rowval, rv0), "bitwise reproducible",
          np.array_equal(nz, Kt2.nzval), "phi rel", abs(phi - phi0) / abs(phi0), "r rel", np.abs(r - r0).max() / np.abs(r0).max())
    err = np.abs(nz - nz0)
    mag = np.maximum(np.abs(nz), np.abs(nz0))
    sc = entry_scales(cp0, rv0, nz0)
    q = err / sc
    print("  err/scale percentiles 50/90/99/99.9/99.99/max", np.percentile(q, [50, 90, 99, 99.9, 99.99, 100]))
    bad = err > 1e-12 * np.maximum(mag, 0.01 * sc)
    print("  violations", int(bad.sum()), "worst ratio", (err / (1e-12 * np.maximum(mag, 0.01 * sc))).max())
    k = np.argsort(q)[-5:]
    col = np.searchsorted(cp0, k + 1, side="right") - 1
    for kk, c in zip(k, col):
        print(f"    entry {kk}: row {rv0[kk]} col {c + 1} got {nz[kk]:.17e} ref {nz0[kk]:.17e} scale {sc[kk]:.3e} err/scale {q[kk]:.2e}")
