"""CPU oracle of the reference's EDG path -- TEST INFRASTRUCTURE ONLY.

Restates `Convection_Diffusion._solveEDG` (reference Python/convection_diffusion.py:42-219) with numpy, element by
element and facet by facet: enriched interior-penalty DG for  -eps Lap u + b . grad u = coeff  with weakly imposed
homogeneous Dirichlet data, plus the inhomogeneous boundary terms of the debug variant (Python/debug_edg.py:191-192,
`assemble(..., boundary_data=True)`).  The product path it checks is csrc/ehdg_edg.cu behind `_solveEDG`; nothing outside
tests/ imports this module.

PARITY STATUS: UNPINNED.  NGSolve holds the arithmetic and is not available (see oracle/ehdg_oracle.py); the
reference stores no EDG matrices or solutions.  What IS pinned: the alpha pencil (same known answers as the EHDG path,
tests/golden/alpha_kat.json) and self-consistency (tests/test_edg_oracle.py: symmetry of the diffusion form, polynomial
exact solutions reproduced, edg == dg when nothing is marked).  Conventions NGSolve decides and the reference does not
show, taken here and tagged for a live check:
  [L] skeleton integrals (`dx(skeleton=True)`) evaluate non-Other() coefficient functions (alpha_stab, mesh_size, the
      normal) on the FIRST element of the facet (facet2el[f, 0]); mesh_size there is |det J| / |e|;
  [M] the compound space is [V, Q_0, Q_1, ...]: V = L2(order) numbered as in SURVEY A.3, Q_i = Compress(L2(order 0)).

Shapes, rules, geometry and marking come from oracle/ehdg_oracle.py.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sps
import scipy.sparse.linalg as spla

from .ehdg_oracle import LOCAL_EDGES, EHDGOracle, Problem, seg_rule, trig_rule


class EDGOracle:
    def __init__(self, verts, tris, el2facet, facet2el, maxh, order, prob: Problem, theta=1e-5):
        self.h = EHDGOracle(verts, tris, el2facet, facet2el, maxh, order, prob)     # marking, geometry, shapes
        self.prob = prob
        self.p = int(order)
        self.theta = float(theta)
        self.ne, self.nf, self.n_p, self.nenr = self.h.ne, self.h.nf, self.h.n_p, self.h.nenr
        self.tris, self.verts = self.h.tris, self.h.verts
        self.el2facet, self.facet2el = self.h.el2facet, self.h.facet2el
        self._numbering()

    # ------------------------------------------------------------------ spaces (:48-56)
    def _numbering(self):
        ne, n_p = self.ne, self.n_p
        off = ne * n_p
        self.offQ, self.q_num = [], []
        for i in range(self.nenr):                                       # mark_dofs(Q, mesh, indicator, size) (:52-54)
            qn = np.full(ne, -1, dtype=np.int64)
            qn[self.h.el_mark[i]] = np.arange(self.h.el_mark[i].sum())
            self.offQ.append(off)
            off += int(self.h.el_mark[i].sum())
            self.q_num.append(qn)
        self.ndof = off
        self.active = np.ones(self.ndof, dtype=bool)                     # ba_active_dofs = fes.FreeDofs(): all (:79-80)

    def el_dofs(self, el):
        d = [self.h.vol_dofs(el)]                                        # offV = 0: same numbers as in the EHDG space
        for i in range(self.nenr):
            q = self.q_num[i][el]
            d.append(np.array([self.offQ[i] + q if q >= 0 else -1]))
        return np.concatenate(d).astype(np.int64)

    @property
    def nloc(self):
        return self.n_p + self.nenr

    # ------------------------------------------------------------------ shapes
    def shapes(self, el, lam):
        """enriched shapes  u() = u[0] + sum u[i+1] psi_i  and their gradients (enrichment_proxy.py:9-79) at local
        barycentrics lam [3, n]: rows [V: n_p][Q_i: 1 each]."""
        phi, gx, gy, dd = self.h.vol_shapes(el, None, lam)
        return phi, gx, gy, dd

    def vol_points(self, order):
        pts, w = trig_rule(order)
        lam = np.stack([pts[:, 0], pts[:, 1], 1.0 - pts[:, 0] - pts[:, 1]])
        return lam, w

    def facet_points(self, el, f, order):
        """points of facet f seen from element el, parametrised from the lower to the higher GLOBAL vertex number so
        that both neighbours visit the same physical points in the same order: (lam [3, n], w |e|, n_out, h_e)."""
        e = int(np.nonzero(self.el2facet[el] == f)[0][0])
        a, b = LOCAL_EDGES[e]
        vn = self.tris[el]
        lo, hi = (a, b) if vn[a] < vn[b] else (b, a)
        t, w = seg_rule(order)
        lam = np.zeros((3, len(t)))
        lam[lo] = 1.0 - t
        lam[hi] = t
        P, J, det = self.h.geom(el)
        A, B = P[a], P[b]
        length = np.linalg.norm(B - A)
        tang = (B - A) / length
        nrm = np.array([tang[1], -tang[0]])
        if np.dot(nrm, A - P[3 - a - b]) < 0:
            nrm = -nrm
        return lam, w * length, nrm, abs(det) / length

    # ------------------------------------------------------------------ pruning (:83-123)
    def prune(self):
        order = 2 * self.p + self.prob.bonus
        lam, w = self.vol_points(order)
        self.factors = {}
        if self.nenr == 0:
            return
        for el in range(self.ne):
            if not self.h.el_marked_any[el]:
                continue
            dofs = self.el_dofs(el)
            N, Nstd = len(dofs), self.n_p
            phi, _, _, _ = self.shapes(el, lam)
            _, _, det = self.h.geom(el)
            elmat = np.einsum("in,jn,n->ij", phi, phi, w * abs(det))     # ipintegrator u() v() (:88-89)
            important = [d >= 0 for d in dofs]
            facs = []
            for i in range(Nstd, N):
                if not important[i]:
                    continue
                act = [j for j in range(i) if important[j]]
                factor = 1 - 2 * sum(elmat[i, j] ** 2 / elmat[i, i] / elmat[j, j] for j in act)
                factor += sum(elmat[i, j] * elmat[i, k] * elmat[j, k] / elmat[i, i] / elmat[j, j] / elmat[k, k]
                              for j in act for k in act)
                factor = np.sqrt(abs(factor))
                facs.append(factor)
                if factor <= self.theta:                                  # :118
                    important[i] = False
                    self.active[dofs[i]] = False
            self.factors[el] = facs

    # ------------------------------------------------------------------ alpha (:129-163)
    def alpha_all(self):
        """generalised eigenvalue of (m, a) on every element, with the reference's carry-over of the previous value when
        the real part is exactly zero (:158-161)."""
        order = 2 * self.p + self.prob.bonus                              # the EDG integrators pass bonus_intorder
        lam_v, w_v = self.vol_points(order)
        out = np.zeros(self.ne)
        csr = 0.0
        for el in range(self.ne):
            _, _, det = self.h.geom(el)
            phi, gx, gy, _ = self.shapes(el, lam_v)
            wd = w_v * abs(det)
            a = np.einsum("in,jn,n->ij", gx, gx, wd) + np.einsum("in,jn,n->ij", gy, gy, wd)
            f = np.einsum("in,n->i", phi, wd) * abs(det) ** (-1.0)        # h ** ((-2 - dim) / 2), volume h = sqrt|det J|
            a = a + np.outer(f, f)
            m = np.zeros_like(a)
            for fc in self.el2facet[el]:
                lam_e, we, nrm, h_e = self.facet_points(el, fc, order)
                _, egx, egy, _ = self.shapes(el, lam_e)
                dn = egx * nrm[0] + egy * nrm[1]
                m += h_e * np.einsum("in,jn,n->ij", dn, dn, we)
            dofs = self.el_dofs(el)
            imp = [i for i, d in enumerate(dofs) if d > 0 and self.active[d]]     # `dof > 0` sic (:148)
            x = np.max(sla.eig(m[np.ix_(imp, imp)], b=a[np.ix_(imp, imp)])[0])
            if x.real:
                csr = float(x.real)
            out[el] = csr
        return out

    # ------------------------------------------------------------------ forms (:165-186) and assembly
    def assemble(self, alpha_stab, boundary_data=False):
        """boundary_data: the right-hand side also carries the weakly imposed Dirichlet data of debug_edg.py:191-192,
        f += exact eps (alpha order^2 / h v - n . grad v) dS + b.n IfPos(b.n, 0, -exact) v dS"""
        p, eps, b = self.p, self.prob.eps, np.asarray(self.prob.beta)
        order = 2 * self.p + self.prob.bonus
        lam_v, w_v = self.vol_points(order)
        rows, cols, vals = [], [], []
        F = np.zeros(self.ndof)

        def add(dr, dc, M):
            keep_r, keep_c = dr >= 0, dc >= 0
            rr, cc = np.meshgrid(dr[keep_r], dc[keep_c], indexing="ij")
            rows.append(rr.ravel()); cols.append(cc.ravel()); vals.append(M[np.ix_(keep_r, keep_c)].ravel())

        for el in range(self.ne):
            _, _, det = self.h.geom(el)
            phi, gx, gy, dd = self.shapes(el, lam_v)
            wd = w_v * abs(det)
            # eps grad u . grad v  -  b u . grad v                 (M[i, j]: test i, trial j)
            M = eps * (np.einsum("in,jn,n->ij", gx, gx, wd) + np.einsum("in,jn,n->ij", gy, gy, wd))
            M -= np.einsum("in,jn,n->ij", b[0] * gx + b[1] * gy, phi, wd)
            d = self.el_dofs(el)
            add(d, d, M)
            fl = np.einsum("in,n->i", phi, wd * self.prob.coeff(dd[0], dd[1]))
            F[d[d >= 0]] += fl[d >= 0]
        for f in range(self.nf):
            e1, e2 = self.facet2el[f]
            lam1, we, n, h1 = self.facet_points(e1, f, order)
            u1, gx1, gy1, _ = self.shapes(e1, lam1)
            dn1 = gx1 * n[0] + gy1 * n[1]
            pen = 4.0 * alpha_stab[e1] / h1 * p ** 2 / h1                 # alpha order^2 / h, alpha = 4 alpha_stab / h
            d1 = self.el_dofs(e1)
            if e2 < 0:
                # dS: eps (pen u v - dn u v - dn v u)
                M = eps * (pen * np.einsum("in,jn,n->ij", u1, u1, we) - np.einsum("in,jn,n->ij", u1, dn1, we)
                           - np.einsum("in,jn,n->ij", dn1, u1, we))
                add(d1, d1, M)
                if boundary_data:
                    _, _, _, dd1 = self.shapes(e1, lam1)
                    gD = self.prob.exact(dd1[0], dd1[1])
                    bn = float(b @ n)
                    fb = eps * np.einsum("in,n->i", pen * u1 - dn1, we * gD)              # exact eps (pen v - dn v)
                    if not bn > 0:
                        fb += -bn * np.einsum("in,n->i", u1, we * gD)                      # b.n IfPos(b.n, 0, -exact) v
                    F[d1[d1 >= 0]] += fb[d1 >= 0]
                continue
            lam2, _, _, _ = self.facet_points(e2, f, order)
            u2, gx2, gy2, _ = self.shapes(e2, lam2)
            dn2 = gx2 * n[0] + gy2 * n[1]                                 # derivative along the FIRST element's normal
            d2 = self.el_dofs(e2)
            bn = float(b @ n)
            # jump = own - other, mean = half sum; signs of the (test, trial) blocks for (1,1), (1,2), (2,1), (2,2)
            jv = {1: (u1, +1.0), 2: (u2, -1.0)}
            dnv = {1: dn1, 2: dn2}
            dofs = {1: d1, 2: d2}
            for ti in (1, 2):
                for tj in (1, 2):
                    vi, si = jv[ti]
                    uj, sj = jv[tj]
                    M = eps * pen * si * sj * np.einsum("in,jn,n->ij", vi, uj, we)
                    M -= eps * 0.5 * si * np.einsum("in,jn,n->ij", vi, dnv[tj], we)      # - {dn u} [v]
                    M -= eps * 0.5 * sj * np.einsum("in,jn,n->ij", dnv[ti], uj, we)      # - {dn v} [u]
                    upwind = (tj == 1) if bn > 0 else (tj == 2)                          # IfPos(b n, u, u.Other())
                    if upwind:
                        M += bn * si * np.einsum("in,jn,n->ij", vi, uj, we)              # b.n u_up [v]
                    add(dofs[ti], dofs[tj], M)
        A = sps.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(self.ndof, self.ndof))
        A.sum_duplicates()
        return A, F

    def solve(self, alpha_stab=None, boundary_data=False):
        """(:189-203) direct solve on the active dofs; inactive entries of the result are zero"""
        if alpha_stab is None:
            alpha_stab = self.alpha_all()
        A, F = self.assemble(alpha_stab, boundary_data=boundary_data)
        act = np.nonzero(self.active)[0]
        x = np.zeros(self.ndof)
        x[act] = spla.spsolve(A[act][:, act].tocsc(), F[act])
        return x

    def l2_error(self, x, order=None):
        """(:205-210) u_h = comp[0] + sum comp[i+1] psi_i ; sqrt(Integrate((u_h - exact)^2, order = 5 + bonus))"""
        order = 5 + self.prob.bonus if order is None else order
        lam, w = self.vol_points(order)
        err2 = 0.0
        for el in range(self.ne):
            _, _, det = self.h.geom(el)
            phi, _, _, dd = self.shapes(el, lam)
            d = self.el_dofs(el)
            coef = np.where(d >= 0, x[np.maximum(d, 0)], 0.0)
            err2 += np.sum(w * abs(det) * (coef @ phi - self.prob.exact(dd[0], dd[1])) ** 2)
        return float(np.sqrt(err2))
