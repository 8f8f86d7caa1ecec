"""ctypes view of the TEST-ONLY host build (tests/hostcheck/hostcheck.cpp) of the element routines."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


class LayerExpr(C.Structure):
    _fields_ = [("c", C.c_double * 4), ("k", C.c_double * 2)]


class CProblem(C.Structure):
    _fields_ = [("order", C.c_int32), ("bonus_int", C.c_int32), ("epsilon", C.c_double),
                ("beta", C.c_double * 2), ("n_enrich", C.c_int32), ("enrich_axis", C.c_int32 * 2),
                ("enrich_k", C.c_double * 2), ("coeff", LayerExpr), ("exact", LayerExpr)]


def c_problem(order, prob) -> CProblem:
    """oracle Problem (with *_desc fields) -> ehdg_problem"""
    cp = CProblem()
    cp.order = order
    cp.bonus_int = prob.bonus
    cp.epsilon = prob.eps
    cp.beta[0], cp.beta[1] = prob.beta
    cp.n_enrich = len(prob.enrich_desc)
    for i, (ax, k) in enumerate(prob.enrich_desc):
        cp.enrich_axis[i] = ax
        cp.enrich_k[i] = k
    for name in ("coeff", "exact"):
        d = getattr(prob, name + "_desc")
        if d is not None:
            for i in range(4):
                getattr(cp, name).c[i] = d[0][i]
            for i in range(2):
                getattr(cp, name).k[i] = d[1][i]
    return cp


def build():
    so = os.path.join(HERE, "libhostcheck.so")
    srcs = [os.path.join(HERE, "hostcheck.cpp")] + [
        os.path.join(ROOT, "convectiondiffusionequation_b200", "csrc", f)
        for f in ("ehdg_general.h", "ehdg_basis.h", "ehdg_rules.h", "ehdg_problem.h", "ehdg_tensors.h", "ehdg_fast.h", "edg_element.h")]
    srcs = [s for s in srcs if os.path.exists(s)]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", so,
                               os.path.join(HERE, "hostcheck.cpp")])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.hc_alpha.restype = C.c_double
        _lib.hc_alpha_bonus.restype = C.c_double
    return _lib


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))
