"""ctypes front-end of oracle/algames_oracle.c (the plain-C restatement used as the timed CPU baseline and as the bulk
parity checker at BASELINE sizes).  TEST / BENCH INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libalgames_oracle.so")
NHIST = 10


def load():
    src = os.path.join(HERE, "algames_oracle.c")
    hdr = os.path.join(HERE, "..", "include", "algames_b200.h")
    if not os.path.exists(LIB) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(LIB):
        subprocess.run(["make", "-s", "-C", HERE], check=True)
    lib = C.CDLL(LIB)
    lib.ago_newton_solve.restype = C.c_int
    lib.ago_newton_solve_ex.restype = C.c_int
    lib.ago_nrow.restype = C.c_int
    return lib


def newton_solve(desc, opts_c, x0, xf, Q, R, uf, Z0, L0, nthreads=0, conlam=None, conmu=None, hist_max=0):
    """desc: algames_b200._capi.ProblemDesc, opts_c: OptionsC; arrays in the ABI layouts.  Returns dict + threads used.
    conlam / conmu [B, N-1, nrow]: multipliers / penalties the solve starts from when opts.dual_reset is false (MPC warm
    start); hist_max > 0 also returns every record!(stats, …) of the solve (same columns as agb_get_history)."""
    lib = load()
    B = x0.shape[0]
    p, N = desc.p, desc.N
    n, m = 4 * p, 2 * p
    nrow = lib.ago_nrow(C.byref(desc))
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (x0, xf, Q, R, uf, Z0, L0)]
    Z = np.empty((B, N, n + m)); L = np.empty((B, p, N - 1, n)); stats = np.empty((B, 10)); status = np.empty(B, dtype=np.int32)
    lam_out = np.empty((B, N - 1, nrow)); mu_out = np.empty((B, N - 1, nrow))
    dp = C.POINTER(C.c_double)
    opt = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    lam0, mu0 = opt(conlam), opt(conmu)
    hist = np.zeros((B, hist_max, NHIST)) if hist_max > 0 else None
    count = np.zeros(B, dtype=np.int32) if hist_max > 0 else None
    ptr = lambda a: None if a is None else a.ctypes.data_as(dp)
    used = lib.ago_newton_solve_ex(C.byref(desc), C.byref(opts_c), B, int(nthreads), *[a.ctypes.data_as(dp) for a in arrs],
                                   ptr(lam0), ptr(mu0), Z.ctypes.data_as(dp), L.ctypes.data_as(dp), ptr(lam_out), ptr(mu_out),
                                   stats.ctypes.data_as(dp), status.ctypes.data_as(C.POINTER(C.c_int)),
                                   ptr(hist), None if count is None else count.ctypes.data_as(C.POINTER(C.c_int)), int(hist_max))
    out = {"Z": Z, "L": L, "stats": stats, "status": status, "conlam": lam_out, "conmu": mu_out}
    if hist is not None:
        out["hist"], out["hist_count"] = hist, count
    return out, used
