"""ctypes front-end of oracle/algames_oracle.c (the plain-C restatement used as the timed CPU baseline).
TEST / BENCH INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libalgames_oracle.so")


def load():
    src = os.path.join(HERE, "algames_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(src) > os.path.getmtime(LIB):
        subprocess.run(["make", "-s", "-C", HERE], check=True)
    lib = C.CDLL(LIB)
    lib.ago_newton_solve.restype = C.c_int
    return lib


def newton_solve(desc, opts_c, x0, xf, Q, R, uf, Z0, L0, nthreads=0):
    """desc: algames_b200._capi.ProblemDesc, opts_c: OptionsC; arrays in the ABI layouts.  Returns dict + threads used."""
    lib = load()
    B = x0.shape[0]
    p, N = desc.p, desc.N
    n, m = 4 * p, 2 * p
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (x0, xf, Q, R, uf, Z0, L0)]
    Z = np.empty((B, N, n + m)); L = np.empty((B, p, N - 1, n)); stats = np.empty((B, 10)); status = np.empty(B, dtype=np.int32)
    dp = C.POINTER(C.c_double)
    used = lib.ago_newton_solve(C.byref(desc), C.byref(opts_c), B, int(nthreads), *[a.ctypes.data_as(dp) for a in arrs],
                                Z.ctypes.data_as(dp), L.ctypes.data_as(dp), stats.ctypes.data_as(dp),
                                status.ctypes.data_as(C.POINTER(C.c_int)))
    return {"Z": Z, "L": L, "stats": stats, "status": status}, used
