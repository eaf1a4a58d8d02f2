"""ctypes binding + helpers for oracle/osqp_ref.c -- TEST / BASELINE INFRASTRUCTURE ONLY (parity unpinned, see
oracle/mpc_oracle.py).  Restates how the reference drives OSQP: one persistent model per thread, `JuMP.fix`
updates of x[:,1] / references between solves (src/main/computation_mpc.jl:17-55)."""
from __future__ import annotations

import ctypes as C
import os
import pathlib
import subprocess
import numpy as np
import scipy.sparse as sp

_HERE = pathlib.Path(__file__).resolve().parent
_LIB = None


class Settings(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("rho", "sigma", "alpha", "eps_abs", "eps_rel", "eps_prim_inf", "eps_dual_inf",
                                          "adaptive_rho_tolerance")] + \
               [(n, C.c_int) for n in ("max_iter", "scaling", "check_termination", "adaptive_rho", "adaptive_rho_interval",
                                       "warm_start", "reset_rho_each_solve")]


def build(force=False):
    so = _HERE / "libosqp_ref.so"
    src = _HERE / "osqp_ref.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
        _LIB.osqp_ref_setup.restype = C.c_void_p
        _LIB.osqp_ref_rho.restype = C.c_double
    return _LIB


def default_settings(**kw) -> Settings:
    s = Settings()
    lib().osqp_ref_default_settings(C.byref(s))
    for k, v in kw.items():
        assert hasattr(s, k), k
        setattr(s, k, v)
    return s


def _ip(a): return np.ascontiguousarray(a, np.int32).ctypes.data_as(C.POINTER(C.c_int))
def _dp(a): return a.ctypes.data_as(C.POINTER(C.c_double))


def min_degree_ordering(K: sp.spmatrix) -> np.ndarray:
    """Plain minimum-degree ordering on the explicit elimination graph (stand-in for OSQP's AMD; exactness of
    the ordering does not affect the iterates, only the fill / speed of the factorisation)."""
    K = sp.csr_matrix(K); n = K.shape[0]
    adj = [set(K.indices[K.indptr[i]:K.indptr[i + 1]]) - {i} for i in range(n)]
    for i in range(n):
        for j in list(adj[i]): adj[j].add(i)
    alive = np.ones(n, bool); order = []
    import heapq
    heap = [(len(adj[i]), i) for i in range(n)]; heapq.heapify(heap)
    while heap:
        d, i = heapq.heappop(heap)
        if not alive[i] or d != len(adj[i]): continue
        alive[i] = False; order.append(i)
        nb = list(adj[i])
        for a in nb:
            adj[a].discard(i)
        for a in nb:
            new = set(nb) - {a} - adj[a]
            if new: adj[a] |= new
        for a in nb: heapq.heappush(heap, (len(adj[a]), a))
        adj[i] = set()
    return np.array(order, np.int32)


def kkt_pattern(P: sp.spmatrix, A: sp.spmatrix):
    n = P.shape[0]; m = A.shape[0]
    return sp.bmat([[sp.csc_matrix(P) + sp.eye(n), A.T], [A, sp.eye(m)]], format="csr")


class Problem:
    """CSC arrays of one OSQP problem  min 1/2 v'Pv + q'v, l <= Av <= u  (P passed as the full symmetric matrix)."""

    def __init__(self, P, q, A, l, u, ordering="mindeg"):
        Pu = sp.triu(sp.csc_matrix(P), format="csc"); Pu.sort_indices()
        Ac = sp.csc_matrix(A); Ac.sort_indices()
        self.n = Pu.shape[0]; self.m = Ac.shape[0]
        self.Pp, self.Pi, self.Px = Pu.indptr.astype(np.int32), Pu.indices.astype(np.int32), Pu.data.astype(float)
        self.Ap, self.Ai, self.Ax = Ac.indptr.astype(np.int32), Ac.indices.astype(np.int32), Ac.data.astype(float)
        self.q = np.ascontiguousarray(q, float); self.l = np.ascontiguousarray(l, float); self.u = np.ascontiguousarray(u, float)
        if ordering == "mindeg":
            self.perm = min_degree_ordering(kkt_pattern(P, A))
        elif ordering == "natural":
            self.perm = np.arange(self.n + self.m, dtype=np.int32)
        else:
            self.perm = np.ascontiguousarray(ordering, np.int32)

    def _args(self):
        return (C.c_int(self.n), C.c_int(self.m), _ip(self.Pp), _ip(self.Pi), _dp(self.Px), _dp(self.q), _ip(self.Ap), _ip(self.Ai),
                _dp(self.Ax), _dp(self.l), _dp(self.u), _ip(self.perm))


def solve_batch(prob: Problem, settings: Settings, upd_rows, upd_vals, sel_cols, cold_start=True, nthreads=0):
    """Solve nb problems that differ only in the l=u value of `upd_rows` (the reference's JuMP.fix updates)."""
    L = lib()
    upd_rows = np.ascontiguousarray(upd_rows, np.int32); upd_vals = np.ascontiguousarray(upd_vals, float)
    nb = upd_vals.shape[0] if upd_vals.ndim == 2 else 1
    k = len(upd_rows); assert upd_vals.size == nb * k
    sel_cols = np.ascontiguousarray(sel_cols, np.int32); ns = len(sel_cols)
    xs = np.zeros((nb, ns)); status = np.zeros(nb, np.int32); iters = np.zeros(nb, np.int32)
    pri = np.zeros(nb); dua = np.zeros(nb); obj = np.zeros(nb)
    rc = L.osqp_ref_solve_batch(*prob._args(), C.byref(settings), C.c_int(nb), C.c_int(k), _ip(upd_rows), _dp(upd_vals),
                                C.c_int(1 if cold_start else 0), C.c_int(ns), _ip(sel_cols), _dp(xs),
                                status.ctypes.data_as(C.POINTER(C.c_int)), iters.ctypes.data_as(C.POINTER(C.c_int)),
                                _dp(pri), _dp(dua), _dp(obj), C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError(f"osqp_ref_solve_batch failed rc={rc}")
    return {"x": xs, "status": status, "iters": iters, "prim_res": pri, "dual_res": dua, "obj": obj}


class Workspace:
    """Single persistent OSQP workspace (what one JuMP model holds); used for the closed-loop latency baseline."""

    def __init__(self, prob: Problem, settings: Settings):
        self.prob = prob; self.L = lib()
        self.w = C.c_void_p(self.L.osqp_ref_setup(*prob._args(), C.byref(settings)))
        if not self.w: raise RuntimeError("osqp_ref_setup failed")

    def update_bounds(self, rows, vals):
        rows = np.ascontiguousarray(rows, np.int32); vals = np.ascontiguousarray(vals, float)
        self.L.osqp_ref_update_bounds(self.w, C.c_int(len(rows)), _ip(rows), _dp(vals), _dp(vals))

    def solve(self, cold_start=False):
        x = np.zeros(self.prob.n); y = np.zeros(self.prob.m)
        it = C.c_int(); pr = C.c_double(); du = C.c_double(); ob = C.c_double(); ru = C.c_int()
        st = self.L.osqp_ref_solve(self.w, C.c_int(1 if cold_start else 0), _dp(x), _dp(y), C.byref(it), C.byref(pr), C.byref(du),
                                   C.byref(ob), C.byref(ru))
        return {"x": x, "y": y, "status": st, "iters": it.value, "prim_res": pr.value, "dual_res": du.value, "obj": ob.value,
                "rho_updates": ru.value, "rho": self.L.osqp_ref_rho(self.w)}

    def kkt_solve(self, b):
        b = np.array(b, float); self.L.osqp_ref_kkt_solve(self.w, _dp(b)); return b

    def nnz_L(self): return self.L.osqp_ref_kkt_nnz_L(self.w)

    def __del__(self):
        try: self.L.osqp_ref_cleanup(self.w)
        except Exception: pass


def max_threads(): return lib().osqp_ref_max_threads()
