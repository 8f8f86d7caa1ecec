"""Generates tests/golden/enriched_solve_oracle.json: L2 / H1 errors of the ENRICHED discrete solution computed by the numpy
oracle (oracle/ehdg_oracle.py) two mathematically equivalent ways -- the reference's uncondensed direct solve
(convection_diffusion.py:385-387, scipy splu for PARDISO) and static condensation + direct facet solve -- on structured
meshes, plus the plain HDG error on the same mesh.  Where the two disagree the reference algorithm has no reproducible
answer (the enriched system is singular to working precision: near-dependent enrichment dofs survive the 1e-5 pruning
test, alpha_K on marked elements is under-integrated); the GPU parity test asserts 1e-8 agreement exactly where they agree
and the CPU suite pins the disagreement at the BASELINE parameters.

    python tests/golden/make_enriched_solve.py          (about 15 minutes, numpy oracle)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from convectiondiffusionequation_b200.mesh import unit_square_mesh   # noqa: E402  (mesh generator only)
from tests.common import make_pair                                    # noqa: E402

# (lattice N, order, eps): resolved-layer cases, then the BASELINE config-2 / config-3 parameters at oracle-reachable sizes
CASES = [(8, 1, 1e-2), (16, 1, 1e-2), (8, 2, 1e-2), (16, 2, 1e-2), (16, 3, 1e-2), (16, 4, 1e-2), (32, 2, 1e-3), (32, 4, 1e-3),
         (16, 3, 1e-4), (32, 3, 1e-4), (16, 4, 1e-6), (32, 4, 1e-6)]


def main():
    out = []
    for (n, p, eps) in CASES:
        mesh = unit_square_mesh(n, kind="structured")
        o, _ = make_pair(mesh, p, eps=eps, enriched=True)
        o.prune()
        lams = o.alpha_all()
        ef = o.l2_error(o.solve_full(lams), h1=True)
        ec = o.l2_error(o.solve_condensed(lams), h1=True)
        o2, _ = make_pair(mesh, p, eps=eps, enriched=False)
        eh = o2.l2_error(o2.solve_condensed(o2.alpha_all()), h1=True)
        rec = dict(n=n, order=p, eps=eps, lam_max=float(lams.max()), l2_full=float(ef[0]), h1_full=float(ef[1]), l2_condensed=float(ec[0]),
                   h1_condensed=float(ec[1]), l2_hdg=float(eh[0]), h1_hdg=float(eh[1]),
                   rel_disagreement=float(max(abs(ef[0] - ec[0]) / min(ef[0], ec[0]), abs(ef[1] - ec[1]) / min(ef[1], ec[1]))))
        print(rec, flush=True)
        out.append(rec)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "enriched_solve_oracle.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
