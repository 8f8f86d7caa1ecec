import sys; sys.path.insert(0,'/root/repo')
from convectiondiffusionequation_b200 import run
from convectiondiffusionequation_b200.convection_diffusion import Convection_Diffusion
for order in (1, 2):
    for ms in (0.25, 0.0625):
        for meth, kw in (("gmres", dict(gmres_restart=60)), ("gmres", dict(gmres_restart=300)), ("dqgmres", dict(gmres_restart=40))):
            cfg = dict(run.config); cfg.update({'order': [order], 'mesh_size': [ms], 'solver': meth, 'gmres_max_iters': 30000}); cfg.update(kw)
            CT = Convection_Diffusion(cfg)
            r, a = CT._solveEHDG()
            print(order, ms, meth, kw, CT.solver_info[-1], float(r['Error'].iloc[0]))
