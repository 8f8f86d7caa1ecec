"""Shared helpers of the parity tests: oracle <-> product problem descriptions."""
from __future__ import annotations

import numpy as np

from oracle import ehdg_oracle as orc


def make_pair(mesh, order, beta=(2.0, 0.001), eps=0.01, bonus=10, enriched=True, literal=False):
    """(oracle object, product ProblemSpec) for the run.py problem on `mesh`.  literal=True switches the product's selection rule
    for dependent volume enrichments off (block-by-block comparisons with the oracle, which keeps every unpruned dof)."""
    from convectiondiffusionequation_b200 import boundary_layer, x, y
    from convectiondiffusionequation_b200.hot_path import ProblemSpec
    prob = orc.run_py_problem(beta=beta, eps=eps, bonus_int=bonus, enriched=enriched)
    o = orc.EHDGOracle(mesh.verts, mesh.tris, mesh.el2facet, mesh.facet2el, mesh.maxh, order, prob)
    p, q = boundary_layer(x, beta[0], eps), boundary_layer(y, beta[1], eps)
    spec = ProblemSpec(order=order, epsilon=eps, beta=beta, bonus_int=bonus, coeff=beta[1] * p + beta[0] * q, exact=p * q,
                       enrich_functions=(p, q) if enriched else (),
                       enrich_domain_ind=(lambda X, Y, h: X > 1 - h, lambda X, Y, h: Y > 1 - h) if enriched else (),
                       vol_drop_tol=0.0 if literal else 1e-6)
    return o, spec


def oracle_masks(o):
    """elem_mask, facet_mask (existence) and elem_active, facet_active (after pruning) as the C ABI defines them."""
    ne, nf = o.ne, o.nf
    em = np.zeros(ne, np.uint8); fm = np.zeros(nf, np.uint8)
    ea = np.zeros(ne, np.uint8); fa = np.zeros(nf, np.uint8)
    for i in range(o.nenr):
        em |= (o.el_mark[i].astype(np.uint8) << i)
        fm |= (o.fac_mark[i].astype(np.uint8) << i)
        q = o.q_num[i]
        act = np.zeros(ne, bool)
        idx = np.nonzero(q >= 0)[0]
        act[idx] = o.active[o.offQ[i] + q[idx]]
        ea |= (act.astype(np.uint8) << i)
        qf = o.qf_num[i]
        actf = np.zeros(nf, bool)
        idx = np.nonzero(qf >= 0)[0]
        actf[idx] = o.active[o.offQF[i] + qf[idx]]
        fa |= (actf.astype(np.uint8) << i)
    return em, fm, ea, fa


def oracle_slot_maps(o):
    """local oracle indices of the padded slot layout: volume slots, facet slots (edge major)."""
    vidx = list(o.loc_V()) + [o.loc_Q(i) for i in range(o.nenr)]
    fidx = []
    for e in range(3):
        fidx += list(o.loc_F(e)) + [o.loc_QF(i, e) for i in range(o.nenr)]
    return np.array(vidx), np.array(fidx)


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)
