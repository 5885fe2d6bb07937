"""Adds "history_exact" to tests/golden/histories.json: after every V-cycle of
the REFERENCE (oracle/_ref, test_mg_3d.c flow) the reference's own
calculateResidual writes the residual field, and its 2-norm is summed here in
long double (chunked), i.e. correctly rounded to double.  The reference's
printed norms ("history") are sequential double sums of the same field and
deviate from this by up to 3e-10 at 513^3 (their rounding, not the field's).

    python oracle/gen_exact_history.py [keys ...]     (default: all up to 513^3)
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import Ref, c_dp  # noqa: E402

PATH = os.path.join(ROOT, "tests", "golden", "histories.json")


def exact_norm(r):
    flat = r.reshape(-1)
    tot = np.longdouble(0)
    step = 1 << 24
    for a in range(0, flat.size, step):
        x = flat[a:a + step].astype(np.longdouble)
        tot += np.sum(x * x)
    return float(np.sqrt(tot))


def main():
    ref = Ref()
    assert ref.available, "build oracle/_ref first (make -C oracle)"
    out = json.load(open(PATH))
    keys = sys.argv[1:] or [k for k in out if out[k]["N"] <= 513]
    L = ref.L
    L.setupBoundaryConditions.argtypes = [c_dp, C.c_int, C.c_double]
    for key in keys:
        g = out[key]
        grid, rhs, h = c_dp(), c_dp(), C.c_double()
        N = L.ref_solver_open(g["coarse"], g["levels"], g["gs"], C.byref(grid), C.byref(rhs),
                              C.byref(h))
        L.SolverSetupBoundaryConditions()
        L.setupBoundaryConditions(grid, N, h.value)
        u = np.ctypeslib.as_array(grid, shape=(N, N, N))
        d = np.ctypeslib.as_array(rhs, shape=(N, N, N))
        r = np.zeros((N, N, N))
        exact = []
        for c in range(g["cycles"]):
            L.ref_vcycle()
            ref.residual(u, d, h.value, r)
            exact.append(exact_norm(r))
        L.ref_solver_close()
        g["history_exact"] = exact
        dev = max(abs(a - b) / b for a, b in zip(g["history"], exact))
        print(key, "max rel deviation of the reference's sequential sums from the exact norm:",
              f"{dev:.2e}", flush=True)
        json.dump(out, open(PATH, "w"), indent=1)


if __name__ == "__main__":
    main()
