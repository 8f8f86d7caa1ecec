"""Driver mirroring the reference's Python/run.py: same problem, same config keys, same two runs
(enriched, then plain HDG), `python -m convectiondiffusionequation_b200.run`.

Known defects of the reference script that are not replicated: run.py:53-54 runs
_solveEHDG twice and calls the first result `edg_table`; run.py:64 references an undefined name.
"""
import pandas as pd

from .coefficients import boundary_layer, x, y
from .convection_diffusion import Convection_Diffusion
from .utils import plot_error_mesh

### Problem configuration (run.py:24-30)
beta = (2, 0.001)
eps = 0.01
p = lambda s: boundary_layer(s, beta[0], eps)
q = lambda s: boundary_layer(s, beta[1], eps)

exact = p(x) * q(y)
coeff = beta[1] * p(x) + beta[0] * q(y)

config = {
    'order': [1, 2],
    'beta': (beta[0], beta[1]),
    'mesh_size': [0.0625],
    'epsilon': 0.01,
    'exact': exact,
    'coeff': coeff,
    'bonus_int': 10,
    'theta': 1e-5,
    'enrich_functions': [p(x), q(y)],
    'enrich_domain_ind': [lambda x, y, h: x > 1 - h, lambda x, y, h: y > 1 - h],
}


def main(write_csv=True, plot=None):
    # with enrichment
    CT = Convection_Diffusion(config)
    ehdg_table, alpha_ehdg = CT._solveEHDG()

    # without enrichment
    plain = dict(config)
    plain.update({'enrich_functions': [], 'enrich_domain_ind': []})
    CT = Convection_Diffusion(plain)
    hdg_table, alpha_hdg = CT._solveEHDG()

    alphas = pd.concat([alpha_ehdg, alpha_hdg])
    if write_csv:
        alphas.to_csv('alphas_dg.csv')
    hdg = pd.concat([hdg_table, ehdg_table])
    print(hdg.to_string(index=False))
    if plot if plot is not None else write_csv:
        print('error plot:', plot_error_mesh(hdg))       # run.py:71 (saved instead of shown: headless)
    return hdg, alphas


if __name__ == "__main__":
    main()
