"""Device-level check of the moment right-hand side: run once with the default library (saves z, g) and once with a build
made with -DEHDG_NO_RHS_MOMENTS (quadrature everywhere; compares).  AB_N / AB_P / AB_KIND as in ab_time.py."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import make_spec
from convectiondiffusionequation_b200.hot_path import EHDGPipeline
from convectiondiffusionequation_b200.mesh import unit_square_mesh

n, order = int(os.environ.get("AB_N", "708")), int(os.environ.get("AB_P", "4"))
mesh = unit_square_mesh(n, kind=os.environ.get("AB_KIND", "structured"))
P = EHDGPipeline(mesh, make_spec(order, 1e-6)).setup().assemble()
torch.cuda.synchronize()
npoly = P.plan_sizes.n_poly
z = P.z[:npoly * P.sizes.n_p].cpu().numpy().reshape(npoly, -1)
g = P.g[:npoly * 3 * P.sizes.nfl].cpu().numpy().reshape(npoly, -1)
path = os.environ.get("MQ_FILE", "/tmp/moment_vs_quadrature.npz")
if os.environ.get("MQ_SAVE") == "1":
    np.savez(path, z=z, g=g, frac=P.moment_rhs_fraction())
    print(json.dumps(dict(saved=path, moment_rhs_fraction=P.moment_rhs_fraction())))
else:
    ref = np.load(path)
    rz = np.linalg.norm(z - ref["z"], axis=1) / np.maximum(np.linalg.norm(ref["z"], axis=1), 1e-300)
    rg = np.linalg.norm(g - ref["g"], axis=1) / np.maximum(np.linalg.norm(ref["g"], axis=1), 1e-300)
    print(json.dumps(dict(lib=os.environ.get("EHDG_LIB_NAME"), elements=int(npoly), moment_rhs_fraction_of_default=float(ref["frac"]),
                          max_rel_diff_z=float(rz.max()), max_rel_diff_g=float(rg.max()),
                          elements_differing=int(((rz > 0) | (rg > 0)).sum()))))
