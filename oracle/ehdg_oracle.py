"""CPU ORACLE (test infrastructure, NOT product code) for the (E)HDG hot path.

PARITY UNPINNED AT MATRIX-ENTRY LEVEL: the arithmetic of the reference lives in the
un-vendored third-party package NGSolve/Netgen (README.md:5 of the reference, version not
pinned, not installable here).  This file restates, in plain numpy and element by element,
  * the reference's own Python logic   (Python/convection_diffusion.py:223-403,
                                        Python/enrichment_proxy.py:83-251, Python/run.py:24-46)
  * NGSolve's published conventions    (Dubiner L2 basis on vertex-sorted
    barycentrics, Legendre facet basis, Duffy Gauss-Jacobi x Gauss-Legendre rules,
    mesh_size semantics, compound-space dof numbering, Compress, FreeDofs).
It is pinned only by the mesh-independent known answers derived from the reference's single
stored artefact Python/alphas_dg.csv (tests/golden/alpha_kat.json, tests/test_oracle_kat.py)
and by its own self-consistency checks (condensed == full direct solve, patch test, rates).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
Nothing under convectiondiffusionequation_b200/ imports it.

Local dof layout of one element (conv_diff.py:246 "[V, F, Qx, QFx, Qy, QFy]"):
    [V : n_p] [F : 3 x (p+1), local edge major] then per enrichment i: [Q_i : 1] [QF_i : 3]
"""
from __future__ import annotations

import functools

import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spla
import scipy.special as spsp

LOCAL_EDGES = ((2, 0), (1, 2), (0, 1))        # NGSolve ET_TRIG edges
REF_VERTS = np.array([[1.0, 0.0], [0.0, 1.0], [0.0, 0.0]])   # V0, V1, V2  (lam0 = x, lam1 = y)


# ----------------------------------------------------------------------------------------------
# quadrature
# ----------------------------------------------------------------------------------------------
def gauss_legendre_01(n):
    x, w = spsp.roots_legendre(n)
    return 0.5 * (x + 1.0), 0.5 * w


def gauss_jacobi10_01(n):
    """nodes/weights for int_0^1 (1-s) f(s) ds."""
    x, w = spsp.roots_jacobi(n, 1.0, 0.0)
    return 0.5 * (x + 1.0), 0.25 * w


def npts_1d(order):
    return int(order) // 2 + 1


def trig_rule(order):
    """Duffy-collapsed tensor rule on the reference triangle, exact to `order`."""
    n = npts_1d(order)
    xi, wx = gauss_jacobi10_01(n)
    eta, we = gauss_legendre_01(n)
    X = np.repeat(xi, n)
    Y = (eta[None, :] * (1.0 - xi[:, None])).reshape(-1)
    W = (wx[:, None] * we[None, :]).reshape(-1)
    return np.stack([X, Y], axis=1), W


def seg_rule(order):
    return gauss_legendre_01(npts_1d(order))


# ----------------------------------------------------------------------------------------------
# bases
# ----------------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def dubiner_coeffs(p):
    """Monomial coefficients C[k, a, b] of phi_k(X, Y) = sum C X^a Y^b, k enumerated i-major.

    phi_ij = P_i^scaled(2Y + X - 1; 1 - X) * P_j^(2i+1,0)(2X - 1)   (NGSolve DubinerBasis::Eval)
    Built symbolically (exact rationals) so it is independent of the recurrences the CUDA
    library uses."""
    import sympy as sp
    X, Y, z = sp.symbols("X Y z")
    out = []
    for i in range(p + 1):
        Pi = sp.Poly(sp.legendre(i, z), z)
        scaled = 0
        for (k,), c in Pi.terms():
            scaled += c * (2 * Y + X - 1) ** k * (1 - X) ** (i - k)
        for j in range(p - i + 1):
            f = sp.expand(scaled * sp.jacobi(j, 2 * i + 1, 0, 2 * X - 1))
            C = np.zeros((p + 1, p + 1))
            for (a, b), c in sp.Poly(f, X, Y).terms():
                C[a, b] = float(c)
            out.append(C)
    return np.array(out)


def dubiner_eval(p, X, Y):
    """values, d/dX, d/dY  of the n_p Dubiner functions at points (X, Y)  -> [n_p, npts] each."""
    C = dubiner_coeffs(p)
    a = np.arange(p + 1)
    Xp = X[None, :] ** a[:, None]
    Yp = Y[None, :] ** a[:, None]
    dXp = np.zeros_like(Xp)
    dYp = np.zeros_like(Yp)
    dXp[1:] = a[1:, None] * Xp[:-1]
    dYp[1:] = a[1:, None] * Yp[:-1]
    v = np.einsum("kab,an,bn->kn", C, Xp, Yp)
    dX = np.einsum("kab,an,bn->kn", C, dXp, Yp)
    dY = np.einsum("kab,an,bn->kn", C, Xp, dYp)
    return v, dX, dY


def legendre_eval(p, xi):
    return np.array([spsp.eval_legendre(k, xi) for k in range(p + 1)])


# ----------------------------------------------------------------------------------------------
class Problem:
    """Problem data of run.py:24-46 as plain numpy callables (independent of the product's
    coefficient classes).  Every callable takes the DISTANCES to the outflow sides,
    (dx, dy) = (1 - x, 1 - y), which the oracle interpolates barycentrically from the vertex
    values 1 - x_v >= 0: the exponent beta (x - 1) / eps = -k dx is then free of cancellation
    (k reaches 2e6 at eps = 1e-6, where forming x first would cost 2e-10 relative accuracy).
    enrich[i] = (psi, dpsi_dx, dpsi_dy), indicators[i] = f(x, y, h) as in run.py:45."""

    def __init__(self, beta, eps, coeff=None, exact=None, enrich=(), indicators=(), bonus_int=10,
                 exact_grad=None, enrich_desc=(), coeff_desc=None, exact_desc=None):
        self.beta = (float(beta[0]), float(beta[1]))
        self.eps = float(eps)
        self.coeff = coeff
        self.exact = exact
        self.exact_grad = exact_grad
        self.enrich = list(enrich)
        self.indicators = list(indicators)
        self.bonus = int(bonus_int)
        self.enrich_desc = list(enrich_desc)      # [(axis, k)]      for the C-ABI under test
        self.coeff_desc = coeff_desc              # (c[4], k[2])     "
        self.exact_desc = exact_desc


def layer_1d(beta, eps):
    """p(s) of run.py:26-27 and its derivative, as functions of d = 1 - s."""
    k = beta / eps
    em = np.exp(-k)

    def val(d):
        return (1.0 - d) + (np.exp(-k * d) - em) / (em - 1.0)

    def der(d):
        return 1.0 + k * np.exp(-k * d) / (em - 1.0)

    return val, der


def run_py_problem(beta=(2.0, 0.001), eps=0.01, bonus_int=10, enriched=True):
    """run.py:24-46 verbatim: exact = p(x) q(y), coeff = beta1 p(x) + beta0 q(y)."""
    p, dp = layer_1d(beta[0], eps)
    q, dq = layer_1d(beta[1], eps)
    zero = lambda dx, dy: np.zeros_like(dx)
    enrich = [(lambda dx, dy: p(dx), lambda dx, dy: dp(dx), zero),
              (lambda dx, dy: q(dy), zero, lambda dx, dy: dq(dy))]
    inds = [lambda x, y, h: x > 1 - h, lambda x, y, h: y > 1 - h]
    kx, ky = beta[0] / eps, beta[1] / eps
    return Problem(beta, eps,
                   coeff=lambda dx, dy: beta[1] * p(dx) + beta[0] * q(dy),
                   exact=lambda dx, dy: p(dx) * q(dy),
                   exact_grad=lambda dx, dy: (dp(dx) * q(dy), p(dx) * dq(dy)),
                   enrich=enrich if enriched else [], indicators=inds if enriched else [],
                   bonus_int=bonus_int,
                   enrich_desc=[(0, kx), (1, ky)] if enriched else [],
                   coeff_desc=((0.0, beta[1], beta[0], 0.0), (kx, ky)),
                   exact_desc=((0.0, 0.0, 0.0, 1.0), (kx, ky)))


# ----------------------------------------------------------------------------------------------
class EHDGOracle:
    def __init__(self, verts, tris, el2facet, facet2el, maxh, order, prob: Problem):
        self.verts = np.asarray(verts, dtype=np.float64)
        self.tris = np.asarray(tris, dtype=np.int64)
        self.el2facet = np.asarray(el2facet, dtype=np.int64)
        self.facet2el = np.asarray(facet2el, dtype=np.int64)
        self.maxh = float(maxh)
        self.p = int(order)
        self.prob = prob
        self.ne = self.tris.shape[0]
        self.nf = self.facet2el.shape[0]
        self.n_p = (self.p + 1) * (self.p + 2) // 2
        self.nenr = len(prob.enrich)
        self.nloc = self.n_p + 3 * (self.p + 1) + 4 * self.nenr
        self._marks()
        self._numbering()

    # ------------------------------------------------------------------ a1: marking
    def _marks(self):
        """enrichment_proxy.py:201-208 (mark_elements), :228-251 (mark_element_bnd)."""
        x, y = self.verts[:, 0], self.verts[:, 1]
        self.el_mark = np.zeros((self.nenr, self.ne), dtype=bool)
        self.fac_mark = np.zeros((self.nenr, self.nf), dtype=bool)
        for i, ind in enumerate(self.prob.indicators):
            vflag = np.array([bool(ind(x[v], y[v], self.maxh)) for v in range(len(x))])
            self.el_mark[i] = vflag[self.tris].any(axis=1)              # any vertex
            for e in np.nonzero(self.el_mark[i])[0]:                    # any adjacent element marked
                self.fac_mark[i, self.el2facet[e]] = True
        self.el_marked_any = self.el_mark.any(axis=0) if self.nenr else np.zeros(self.ne, bool)

    # ------------------------------------------------------------------ dof numbering (A.3-A.5)
    def _numbering(self):
        p, n_p, ne, nf = self.p, self.n_p, self.ne, self.nf
        off = 0
        self.offV = off; off += ne * n_p
        self.offF = off; off += nf * (p + 1)
        self.offQ, self.offQF, self.q_num, self.qf_num = [], [], [], []
        for i in range(self.nenr):
            qn = np.full(ne, -1, dtype=np.int64)                        # Compress(Q, active)
            qn[self.el_mark[i]] = np.arange(self.el_mark[i].sum())
            self.offQ.append(off); off += int(self.el_mark[i].sum())
            fn = np.full(nf, -1, dtype=np.int64)                        # Compress(QF, active)
            fn[self.fac_mark[i]] = np.arange(self.fac_mark[i].sum())
            self.offQF.append(off); off += int(self.fac_mark[i].sum())
            self.q_num.append(qn); self.qf_num.append(fn)
        self.ndof = off
        bnd = self.facet2el[:, 1] < 0
        free = np.ones(self.ndof, dtype=bool)                           # FreeDofs, dirichlet=".*"
        for f in np.nonzero(bnd)[0]:
            free[self.facet_dofs(f)] = False
            for i in range(self.nenr):
                if self.qf_num[i][f] >= 0:
                    free[self.offQF[i] + self.qf_num[i][f]] = False
        self.free = free
        self.active = free.copy()                                       # ba_active_dofs (:263-264)

    def vol_dofs(self, el):
        n_p, ne = self.n_p, self.ne
        return np.concatenate([[self.offV + el],
                               self.offV + ne + el * (n_p - 1) + np.arange(n_p - 1)]).astype(np.int64)

    def facet_dofs(self, f):
        p, nf = self.p, self.nf
        return np.concatenate([[self.offF + f],
                               self.offF + nf + f * p + np.arange(p)]).astype(np.int64)

    def el_dofs(self, el):
        d = [self.vol_dofs(el)]
        for e in range(3):
            d.append(self.facet_dofs(self.el2facet[el, e]))
        for i in range(self.nenr):
            q = self.q_num[i][el]
            d.append(np.array([self.offQ[i] + q if q >= 0 else -1]))
            qf = self.qf_num[i][self.el2facet[el]]
            d.append(np.where(qf >= 0, self.offQF[i] + qf, -1))
        return np.concatenate(d).astype(np.int64)

    # local index helpers
    def loc_V(self):
        return np.arange(self.n_p)

    def loc_F(self, e):
        return self.n_p + e * (self.p + 1) + np.arange(self.p + 1)

    def loc_Q(self, i):
        return self.n_p + 3 * (self.p + 1) + 4 * i

    def loc_QF(self, i, e):
        return self.n_p + 3 * (self.p + 1) + 4 * i + 1 + e

    def loc_voltype(self):
        return np.concatenate([self.loc_V(), [self.loc_Q(i) for i in range(self.nenr)]]).astype(int)

    def loc_factype(self, e):
        return np.concatenate([self.loc_F(e), [self.loc_QF(i, e) for i in range(self.nenr)]]).astype(int)

    # ------------------------------------------------------------------ geometry & shapes
    def geom(self, el):
        P = self.verts[self.tris[el]]
        J = np.stack([P[0] - P[2], P[1] - P[2]], axis=1)               # x = P2 + J (xh, yh)
        det = np.linalg.det(J)
        return P, J, det

    def vol_shapes(self, el, pts, lam=None):
        """enriched volume shapes at reference points: phi, dphi/dx, dphi/dy  [n_p+nenr, npts],
        plus the distances (1-x, 1-y) of the points.  (enrichment_proxy.py:95-96, 133-138)"""
        P, J, det = self.geom(el)
        if lam is None:
            lam = np.stack([pts[:, 0], pts[:, 1], 1.0 - pts[:, 0] - pts[:, 1]])
        vn = self.tris[el]
        f = np.argsort(vn, kind="stable")                               # ascending vertex numbers
        grad_lam = np.array([[1.0, 0.0], [0.0, 1.0], [-1.0, -1.0]])
        v, dX, dY = dubiner_eval(self.p, lam[f[0]], lam[f[1]])
        dref = dX[:, None, :] * grad_lam[f[0]][None, :, None] + dY[:, None, :] * grad_lam[f[1]][None, :, None]
        Jinv = np.linalg.inv(J)
        dphys = np.einsum("ac,kan->kcn", Jinv, dref)                    # J^{-T} grad_ref
        dd = self.dist(el, lam)
        phi = [v]; gx = [dphys[:, 0]]; gy = [dphys[:, 1]]
        for (psi, dpx, dpy) in self.prob.enrich:
            phi.append(psi(dd[0], dd[1])[None]); gx.append(dpx(dd[0], dd[1])[None])
            gy.append(dpy(dd[0], dd[1])[None])
        return np.concatenate(phi), np.concatenate(gx), np.concatenate(gy), dd

    def dist(self, el, lam):
        """(1 - x, 1 - y) at points given by local barycentrics lam [3, npts]."""
        P = self.verts[self.tris[el]]
        return np.stack([lam.T @ (1.0 - P[:, 0]), lam.T @ (1.0 - P[:, 1])])

    def edge_data(self, el, e, order):
        """points on local edge e: reference pts, weights*|e|, unit outward normal, h_e, mu."""
        P, J, det = self.geom(el)
        a, b = LOCAL_EDGES[e]
        t, w = seg_rule(order)
        pts = (1.0 - t)[:, None] * REF_VERTS[a][None] + t[:, None] * REF_VERTS[b][None]
        lam = np.zeros((3, len(t)))                                      # exact barycentrics: the vertex
        lam[a] = 1.0 - t                                                 # opposite the edge gets a true 0
        lam[b] = t
        A, B = P[a], P[b]
        length = np.linalg.norm(B - A)
        tang = (B - A) / length
        nrm = np.array([tang[1], -tang[0]])
        c = P[3 - a - b]
        if np.dot(nrm, A - c) < 0:
            nrm = -nrm
        h_e = abs(det) / length                                         # SURVEY A.2 [V]
        vn = self.tris[el]
        lo, hi = (a, b) if vn[a] < vn[b] else (b, a)                    # sorted edge (A.4)
        xi = lam[hi] - lam[lo]
        mu = legendre_eval(self.p, xi)
        self._edge_lam = lam
        return pts, w * length, nrm, h_e, mu

    def facet_shapes(self, el, e, pts, mu):
        """enriched facet shapes on edge e  [p+1+nenr, npts]  (enrichment_proxy.py:161-162)."""
        dd = self.dist(el, self._edge_lam)
        rows = [mu] + [psi(dd[0], dd[1])[None] for (psi, _, _) in self.prob.enrich]
        return np.concatenate(rows)

    # ------------------------------------------------------------------ a3: pruning
    def boundary_gram(self, el):
        """ipintegrator of conv_diff.py:272-273: oint (u v + uhat vhat), bonus quadrature."""
        order = 2 * self.p + self.prob.bonus
        G = np.zeros((self.nloc, self.nloc))
        iv = self.loc_voltype()
        for e in range(3):
            pts, w, nrm, h_e, mu = self.edge_data(el, e, order)
            phi, _, _, _ = self.vol_shapes(el, pts, self._edge_lam)
            fs = self.facet_shapes(el, e, pts, mu)
            jf = self.loc_factype(e)
            G[np.ix_(iv, iv)] += np.einsum("in,jn,n->ij", phi, phi, w)
            G[np.ix_(jf, jf)] += np.einsum("in,jn,n->ij", fs, fs, w)
        return G

    def prune_element(self, el, threshold=1e-5):
        """conv_diff.py:286-310 for ONE marked element, literally (python floats => ZeroDivisionError aborts the element).
        The loop only looks at the element's own dof list (`important` starts from dofs[i] >= 0), so its decisions do not
        depend on what other elements removed.  Returns (dropped global dofs, [(local index, factor)])."""
        dofs = self.el_dofs(el)
        N = len(dofs)
        elmat = self.boundary_gram(el).tolist()
        important = [bool(dofs[i] >= 0) for i in range(N)]
        factors, dropped = [], []
        try:
            for i in range(self.n_p, N):
                if important[i]:
                    if elmat[i][i] == 0.0:
                        # psi vanishes identically on this facet (the outflow edge itself, d = 0 in the
                        # cancellation-free form used here).  NGSolve evaluates psi from a rounded x and
                        # gets ~1e-16 instead of 0, so the reference neither divides by zero nor prunes
                        # this (Dirichlet) dof: skip it, and keep it out of the later `active` lists.
                        important[i] = False
                        factors.append((i, -3.0))
                        continue
                    act = [j for j in range(i) if important[j]]
                    factor = 1 - 2 * sum([elmat[i][j] ** 2 / elmat[i][i] / elmat[j][j] for j in act])
                    factor += sum([elmat[i][j] * elmat[i][k] * elmat[j][k] / elmat[i][i]
                                   / elmat[j][j] / elmat[k][k] for j in act for k in act])
                    factor = np.sqrt(abs(factor))
                    factors.append((i, factor))
                    if factor <= threshold:
                        important[i] = False
                        if dofs[i] >= 0:
                            dropped.append(int(dofs[i]))
        except ZeroDivisionError:
            pass
        return dropped, factors

    def prune(self, threshold=1e-5):
        """conv_diff.py:282-310: every marked element in turn; a global dof is deactivated as soon as one element says so."""
        self.prune_factors = {}
        if self.nenr == 0:
            return self.active
        for el in range(self.ne):
            if not self.el_marked_any[el]:
                continue
            dropped, factors = self.prune_element(el, threshold)
            for d in dropped:
                self.active[d] = False
            self.prune_factors[el] = factors
        return self.active

    # ------------------------------------------------------------------ a4: stabilisation
    def alpha_matrices(self, el):
        """a_diff (:315), f (:317), m (:320): NO bonus order; full local size, facet rows zero."""
        order = 2 * self.p
        N = self.nloc
        iv = self.loc_voltype()
        P, J, det = self.geom(el)
        pts, w = trig_rule(order)
        phi, gx, gy, _ = self.vol_shapes(el, pts)
        wd = w * abs(det)
        a = np.zeros((N, N)); m = np.zeros((N, N)); f = np.zeros(N)
        a[np.ix_(iv, iv)] = np.einsum("in,jn,n->ij", gx, gx, wd) + np.einsum("in,jn,n->ij", gy, gy, wd)
        hvol = np.sqrt(abs(det))                                         # volume mesh_size (A.2)
        f[iv] = np.einsum("in,n->i", phi, wd) * hvol ** (-2.0)           # h**((-2-dim)/2), dim = 2
        for e in range(3):
            epts, ew, nrm, h_e, mu = self.edge_data(el, e, order)
            _, egx, egy, _ = self.vol_shapes(el, epts, self._edge_lam)
            dn = egx * nrm[0] + egy * nrm[1]
            m[np.ix_(iv, iv)] += h_e * np.einsum("in,jn,n->ij", dn, dn, ew)
        return a, f, m

    def alpha_lambda(self, el):
        """conv_diff.py:322-351 for one element: lambda = max eig(pinv(a + f f^T) m).real."""
        a, f, m = self.alpha_matrices(el)
        a = a + np.outer(f, f)                                           # :329-331
        dofs = self.el_dofs(el)
        important = [i for i, d in enumerate(dofs) if d > 0 and self.active[d]]   # :336-339 (dof > 0 sic)
        mm = m[np.ix_(important, important)]
        aa = a[np.ix_(important, important)]
        x = np.max(np.linalg.eig(np.linalg.pinv(aa) @ mm)[0])            # :345
        return float(np.real(x))

    def alpha_all(self):
        return np.array([self.alpha_lambda(el) for el in range(self.ne)])

    # ------------------------------------------------------------------ a5 + a7: element blocks
    def element_matrix(self, el, lam):
        """Full local matrix (NGSolve layout) of eps*diffusion + convection (:357-375) and the
        load vector (:380-381).  lam = lambda_K of a4; alpha_K = 20 lam (:349)."""
        p, eps = self.p, self.prob.eps
        b = np.array(self.prob.beta)
        order = 2 * p + self.prob.bonus
        N = self.nloc
        iv = self.loc_voltype()
        A = np.zeros((N, N)); F = np.zeros(N)
        P, J, det = self.geom(el)
        pts, w = trig_rule(order)
        phi, gx, gy, dd = self.vol_shapes(el, pts)
        wd = w * abs(det)
        bg = b[0] * gx + b[1] * gy
        A[np.ix_(iv, iv)] += eps * (np.einsum("in,jn,n->ij", gx, gx, wd) + np.einsum("in,jn,n->ij", gy, gy, wd))
        A[np.ix_(iv, iv)] -= np.einsum("in,jn,n->ij", bg, phi, wd)       # -(b u).grad v, rows = test
        if self.prob.coeff is not None:
            F[iv] = np.einsum("in,n->i", phi, wd * self.prob.coeff(dd[0], dd[1]))
        alpha = 20.0 * lam
        for e in range(3):
            epts, ew, nrm, h_e, mu = self.edge_data(el, e, order)
            ephi, egx, egy, _ = self.vol_shapes(el, epts, self._edge_lam)
            fs = self.facet_shapes(el, e, epts, mu)
            jf = self.loc_factype(e)
            dn = egx * nrm[0] + egy * nrm[1]
            bn = float(b @ nrm)
            pos = 1.0 if bn > 0 else 0.0                                  # IfPos(c,a,b): c > 0
            neg = 1.0 - pos
            tau = eps * alpha * p ** 2 / h_e
            M_vv = np.einsum("in,jn,n->ij", ephi, ephi, ew)
            M_vf = np.einsum("in,kn,n->ik", ephi, fs, ew)
            M_ff = np.einsum("kn,ln,n->kl", fs, fs, ew)
            D_vv = np.einsum("in,jn,n->ij", ephi, dn, ew)                 # phi_i dn(phi_j)
            D_fv = np.einsum("kn,jn,n->kj", fs, dn, ew)                   # mu_k dn(phi_j)
            A[np.ix_(iv, iv)] += (tau + pos * bn) * M_vv - eps * D_vv - eps * D_vv.T
            A[np.ix_(iv, jf)] += (-tau + neg * bn) * M_vf + eps * D_fv.T
            A[np.ix_(jf, iv)] += (-tau - pos * bn - pos) * M_vf.T + eps * D_fv
            A[np.ix_(jf, jf)] += (tau - neg * bn + pos) * M_ff
        return A, F

    def local_active(self, el):
        """(volume-type local idx kept, facet-type local idx kept, their global dofs)."""
        dofs = self.el_dofs(el)
        ok = np.array([d >= 0 and self.active[d] for d in dofs])
        iv = [i for i in self.loc_voltype() if ok[i]]
        jf = [j for e in range(3) for j in self.loc_factype(e) if ok[j]]
        jf = sorted(jf)
        return np.array(iv, dtype=int), np.array(jf, dtype=int), dofs

    def condense(self, el, lam):
        """a6: S_K = A_FF - A_FV A_VV^-1 A_VF, g_K = -A_FV A_VV^-1 f_V on active dofs.
        (NGSolve equivalent: BilinearForm(fes, condense=True); not present in the reference.)"""
        A, F = self.element_matrix(el, lam)
        iv, jf, dofs = self.local_active(el)
        Avv = A[np.ix_(iv, iv)]; Avf = A[np.ix_(iv, jf)]
        Afv = A[np.ix_(jf, iv)]; Aff = A[np.ix_(jf, jf)]
        W = np.linalg.solve(Avv, Avf)
        z = np.linalg.solve(Avv, F[iv])
        S = Aff - Afv @ W
        g = -Afv @ z
        return dict(S=S, g=g, W=W, z=z, iv=iv, jf=jf, dofs=dofs, A=A, F=F)

    # ------------------------------------------------------------------ a8/a9: global systems
    def assemble_full(self, lams):
        """acd.Assemble() (:374-377) + f.Assemble() (:380-383): graph over regular dofs, columns
        ascending, structural zeros kept (A.8)."""
        rows, cols, vals = [], [], []
        rhs = np.zeros(self.ndof)
        for el in range(self.ne):
            A, F = self.element_matrix(el, lams[el])
            dofs = self.el_dofs(el)
            reg = np.nonzero(dofs >= 0)[0]
            d = dofs[reg]
            rows.append(np.repeat(d, len(d))); cols.append(np.tile(d, len(d)))
            vals.append(A[np.ix_(reg, reg)].reshape(-1))
            np.add.at(rhs, d, F[reg])
        M = sps.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                           shape=(self.ndof, self.ndof)).tocsr()
        M.sort_indices()
        return M, rhs

    def solve_full(self, lams):
        """:385-387  Inverse(ba_active_dofs, "pardiso") * f  (scipy splu stands in for PARDISO)."""
        M, rhs = self.assemble_full(lams)
        idx = np.nonzero(self.active)[0]
        sub = M[idx][:, idx].tocsc()
        x = np.zeros(self.ndof)
        x[idx] = spla.splu(sub).solve(rhs[idx])
        return x

    def assemble_condensed(self, lams, keep=None):
        """Condensed facet system over the global dof numbering (rows of V dofs empty).
        `keep` = dof mask that defines the graph (default: active dofs)."""
        keep = self.active if keep is None else keep
        rows, cols, vals = [], [], []
        rhs = np.zeros(self.ndof)
        self._cond = []
        for el in range(self.ne):
            c = self.condense(el, lams[el])
            self._cond.append(c)
            d = c["dofs"][c["jf"]]
            rows.append(np.repeat(d, len(d))); cols.append(np.tile(d, len(d)))
            vals.append(c["S"].reshape(-1))
            np.add.at(rhs, d, c["g"])
        S = sps.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                           shape=(self.ndof, self.ndof)).tocsr()
        S.sort_indices()
        return S, rhs

    def solve_condensed(self, lams):
        S, rhs = self.assemble_condensed(lams)
        fd = np.nonzero(np.diff(S.indptr) > 0)[0]
        x = np.zeros(self.ndof)
        x[fd] = spla.splu(S[fd][:, fd].tocsc()).solve(rhs[fd])
        for el, c in enumerate(self._cond):                              # back-solve
            d = c["dofs"]
            x[d[c["iv"]]] = c["z"] - c["W"] @ x[d[c["jf"]]]
        return x

    # ------------------------------------------------------------------ inhomogeneous Dirichlet data (debug_ehdg.py:210-214)
    def dirichlet_values(self):
        """gfu.components[1].Set(exact, BND): L2 projection of `exact` onto the Legendre modes of every boundary facet
        (bonus rule; facet parameter from the lower to the higher global vertex).  {facet: coefficients [p+1]}"""
        t, w = seg_rule(2 * self.p + self.prob.bonus)
        mu = legendre_eval(self.p, 2.0 * t - 1.0)
        out = {}
        for f in np.nonzero(self.facet2el[:, 1] < 0)[0]:
            el = self.facet2el[f, 0]
            e = int(np.nonzero(self.el2facet[el] == f)[0][0])
            a, b = LOCAL_EDGES[e]
            va, vb = self.tris[el][a], self.tris[el][b]
            lo, hi = (va, vb) if va < vb else (vb, va)
            dlo, dhi = 1.0 - self.verts[lo], 1.0 - self.verts[hi]
            g = self.prob.exact((1.0 - t) * dlo[0] + t * dhi[0], (1.0 - t) * dlo[1] + t * dhi[1])
            out[int(f)] = (2.0 * np.arange(self.p + 1) + 1.0) * (mu * (g * w)[None, :]).sum(axis=1)
        return out

    def solve_full_lifted(self, lams):
        """debug_ehdg.py:210-214 literally: gfu = boundary projection; f -= A gfu; gfu += Inverse(active) f"""
        M, rhs = self.assemble_full(lams)
        x0 = np.zeros(self.ndof)
        for f, c in self.dirichlet_values().items():
            x0[self.facet_dofs(f)] = c
        rhs = rhs - M @ x0
        idx = np.nonzero(self.active)[0]
        x = x0.copy()
        x[idx] += spla.splu(M[idx][:, idx].tocsc()).solve(rhs[idx])
        return x

    def solve_condensed_lifted(self, lams):
        """the same through static condensation: the prescribed dofs of boundary facets move to the right-hand side of the
        facet system and return in the element back-solve"""
        ub = self.dirichlet_values()
        x = np.zeros(self.ndof)
        for f, c in ub.items():
            x[self.facet_dofs(f)] = c
        rows, cols, vals = [], [], []
        rhs = np.zeros(self.ndof)
        cond = []
        for el in range(self.ne):
            A, F = self.element_matrix(el, lams[el])
            iv, jf, dofs = self.local_active(el)
            jb = [j for e in range(3) if self.facet2el[self.el2facet[el, e], 1] < 0 for j in self.loc_F(e)]
            jall = np.array(list(jf) + jb, dtype=int)
            W = np.linalg.solve(A[np.ix_(iv, iv)], A[np.ix_(iv, jall)])
            z = np.linalg.solve(A[np.ix_(iv, iv)], F[iv])
            S = A[np.ix_(jall, jall)] - A[np.ix_(jall, iv)] @ W
            g = -A[np.ix_(jall, iv)] @ z
            nfree = len(jf)
            d = dofs[jf]
            rows.append(np.repeat(d, nfree)); cols.append(np.tile(d, nfree)); vals.append(S[:nfree, :nfree].reshape(-1))
            np.add.at(rhs, d, g[:nfree] - S[:nfree, nfree:] @ x[dofs[jb]])
            cond.append((iv, jall, dofs, W, z))
        Sg = sps.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(self.ndof, self.ndof)).tocsr()
        fd = np.nonzero(np.diff(Sg.indptr) > 0)[0]
        x[fd] = spla.splu(Sg[fd][:, fd].tocsc()).solve(rhs[fd])
        for iv, jall, dofs, W, z in cond:
            x[dofs[iv]] = z - W @ x[dofs[jall]]
        return x

    # ------------------------------------------------------------------ a10: recombine + error
    def l2_error(self, x, order=None, h1=False):
        """:389-394  u_h = comp[0] + sum comp[2i+2] psi_i ; sqrt(Integrate((u_h-exact)^2, order=50+bonus))."""
        order = 50 + self.prob.bonus if order is None else order
        pts, w = trig_rule(order)
        err2 = 0.0
        h1e2 = 0.0
        iv = self.loc_voltype()
        for el in range(self.ne):
            P, J, det = self.geom(el)
            phi, gx, gy, dd = self.vol_shapes(el, pts)
            dofs = self.el_dofs(el)[iv]
            coef = np.where(dofs >= 0, x[np.maximum(dofs, 0)], 0.0)
            uh = coef @ phi
            ex = self.prob.exact(dd[0], dd[1])
            err2 += np.sum(w * abs(det) * (uh - ex) ** 2)
            if h1:
                ex_gx, ex_gy = self.prob.exact_grad(dd[0], dd[1])
                h1e2 += np.sum(w * abs(det) * ((coef @ gx - ex_gx) ** 2 + (coef @ gy - ex_gy) ** 2))
        return (np.sqrt(err2), np.sqrt(h1e2)) if h1 else np.sqrt(err2)
