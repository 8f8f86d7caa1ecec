"""CPU ORACLE, fast leg (test infrastructure, NOT product code): ctypes wrapper of oracle/ehdg_oracle.c,
the plain-C + OpenMP restatement of the reference's quadrature algorithm.  Basis tables and rules
are produced by the numpy oracle (oracle/ehdg_oracle.py).  Used by tests/ for parity at sizes the
numpy oracle cannot reach and by bench.py for the CPU baseline / `--impl reference` arm.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

from . import ehdg_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ehdg_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")


def _cpu_tag() -> str:
    """the library is built -O3 -march=native, so it is keyed by the host CPU's feature flags: a copy built in another
    container is not loaded on a different CPU (it is rebuilt there)"""
    import hashlib
    try:
        txt = open("/proc/cpuinfo").read()
        flags = next((ln for ln in txt.splitlines() if ln.startswith("flags")), "")
    except OSError:
        flags = ""
    return hashlib.sha256(flags.encode()).hexdigest()[:10]


LIB = os.path.join(OUT_DIR, f"libehdg_oracle_{_cpu_tag()}.so")


class Params(C.Structure):
    _fields_ = [("p", C.c_int), ("n_p", C.c_int), ("nenr", C.c_int), ("bonus", C.c_int),
                ("eps", C.c_double), ("bx", C.c_double), ("by", C.c_double),
                ("enr_axis", C.c_int * 2), ("enr_k", C.c_double * 2),
                ("coeff_c", C.c_double * 4), ("coeff_k", C.c_double * 2),
                ("mono", C.c_void_p), ("nq", C.c_int), ("nqf", C.c_int), ("nq0", C.c_int), ("nqf0", C.c_int),
                ("vx", C.c_void_p), ("vy", C.c_void_p), ("vw", C.c_void_p), ("st", C.c_void_p), ("sw", C.c_void_p),
                ("vx0", C.c_void_p), ("vy0", C.c_void_p), ("vw0", C.c_void_p), ("st0", C.c_void_p), ("sw0", C.c_void_p)]


def build(force=False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(SRC) > os.path.getmtime(LIB):
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_batch.restype = C.c_int
        _lib.orc_max_threads.restype = C.c_int
    return _lib


class COracle:
    """order/problem specific tables + the batch call"""

    def __init__(self, order, prob: orc.Problem):
        self.p = order
        self.prob = prob
        self.n_p = (order + 1) * (order + 2) // 2
        self.nenr = len(prob.enrich_desc)
        self.nb = order + 1 + self.nenr
        self.nf = 3 * self.nb
        self.nv = self.n_p + self.nenr
        self._keep = []
        P = Params()
        P.p, P.n_p, P.nenr, P.bonus = order, self.n_p, self.nenr, prob.bonus
        P.eps, P.bx, P.by = prob.eps, prob.beta[0], prob.beta[1]
        for i, (ax, k) in enumerate(prob.enrich_desc):
            P.enr_axis[i], P.enr_k[i] = ax, k
        if prob.coeff_desc is not None:
            for i in range(4):
                P.coeff_c[i] = prob.coeff_desc[0][i]
            for i in range(2):
                P.coeff_k[i] = prob.coeff_desc[1][i]

        def hold(a):
            a = np.ascontiguousarray(a, dtype=np.float64)
            self._keep.append(a)
            return a.ctypes.data

        P.mono = hold(orc.dubiner_coeffs(order))
        for suffix, ordr in (("", 2 * order + prob.bonus), ("0", 2 * order)):
            pts, w = orc.trig_rule(ordr)
            t, ws = orc.seg_rule(ordr)
            setattr(P, "nq" + suffix, len(w))
            setattr(P, "nqf" + suffix, len(ws))
            setattr(P, "vx" + suffix, hold(pts[:, 0]))
            setattr(P, "vy" + suffix, hold(pts[:, 1]))
            setattr(P, "vw" + suffix, hold(w))
            setattr(P, "st" + suffix, hold(t))
            setattr(P, "sw" + suffix, hold(ws))
        self.P = P

    def batch(self, verts, tris, el_list, volmask, facmask, lam=None, want_S=True, want_full=False, condense=True, nthreads=0):
        """lam None -> computed.  Returns dict(lam, S [n,nf,nf], g [n,nf], Afull)."""
        el_list = np.ascontiguousarray(el_list, dtype=np.int32)
        n = len(el_list)
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        tris = np.ascontiguousarray(tris, dtype=np.int32)
        volmask = np.ascontiguousarray(volmask, dtype=np.uint8)
        facmask = np.ascontiguousarray(facmask, dtype=np.uint8).reshape(n, 3)
        compute_alpha = lam is None
        lam = np.zeros(n) if lam is None else np.ascontiguousarray(lam, dtype=np.float64).copy()
        S = np.zeros((n, self.nf, self.nf)) if (want_S and condense) else None
        g = np.zeros((n, self.nf)) if (want_S and condense) else None
        N = self.nv + self.nf
        Af = np.zeros((n, N * N + N)) if want_full else None
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        bad = lib().orc_batch(C.byref(self.P), ptr(verts), ptr(tris), n, ptr(el_list), ptr(volmask), ptr(facmask), int(compute_alpha),
                              ptr(lam), ptr(S), ptr(g), ptr(Af), int(condense), int(nthreads))
        return dict(lam=lam, S=S, g=g, Afull=Af, bad=bad)


def marks(mesh, prob):
    """vectorised mark_elements / mark_element_bnd (enrichment_proxy.py:201-251): u8 bit masks"""
    em = np.zeros(mesh.ne, np.uint8)
    fm = np.zeros(mesh.nf, np.uint8)
    x, y = mesh.verts[:, 0], mesh.verts[:, 1]
    for i, ind in enumerate(prob.indicators):
        vflag = np.asarray(ind(x, y, mesh.maxh)).astype(bool)
        e = vflag[mesh.tris].any(axis=1)
        em |= (e.astype(np.uint8) << i)
        f = np.zeros(mesh.nf, bool)
        f[mesh.el2facet[e].reshape(-1)] = True
        fm |= (f.astype(np.uint8) << i)
    return em, fm


def time_sample(n, kind, order, eps, beta, budget_s=20.0, max_elems=None, nthreads=0):
    """Times alpha + element blocks + condensation on an evenly strided sample of the named workload
    on the host cores.  Returns elements/s and a description of the sample."""
    from convectiondiffusionequation_b200.mesh import unit_square_mesh      # mesh generator only (no kernels)
    mesh = unit_square_mesh(n, kind=kind)
    prob = orc.run_py_problem(beta=beta, eps=eps)
    co = COracle(order, prob)
    em, fm = marks(mesh, prob)
    # explicit thread count: under torchrun OMP_NUM_THREADS is 1 and omp_get_max_threads() would follow it
    threads = (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)) if nthreads <= 0 else nthreads

    def run(m):
        m = min(m, mesh.ne)
        els = np.unique(np.linspace(0, mesh.ne - 1, m).astype(np.int32))
        t0 = time.perf_counter()
        co.batch(mesh.verts, mesh.tris, els, em[els], fm[mesh.el2facet[els]], want_S=False, condense=True, nthreads=threads)
        return len(els), time.perf_counter() - t0

    m, dt = run(512)                                  # calibration
    rate = m / dt
    m2 = int(max(512, rate * budget_s * 0.8))
    if max_elems:
        m2 = min(m2, max_elems)
    m2, dt2 = run(m2)
    return {"value": m2 / dt2, "sample_elements": m2, "seconds": dt2, "cores": int(threads), "kind": "port",
            "sample": f"{m2} elements evenly strided over the {mesh.ne}-element {kind} mesh (order {order}, eps {eps}); "
                      f"alpha + quadrature blocks + condensation per element, C/OpenMP oracle, {threads} threads; "
                      "enrichment marks without pruning"}
