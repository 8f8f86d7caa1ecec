"""Derives the mesh-independent known answers held by the reference's only stored artefact,
Python/alphas_dg.csv (per-element lambda_max written by debug_ehdg.py:162-169), and stores them
as tests/golden/alpha_kat.json.  Run in the build container (needs /root/reference):

    python tests/golden/make_alpha_kat.py
"""
import json
import os

import pandas as pd

SRC = "/root/reference/Python/alphas_dg.csv"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "alpha_kat.json")


def main():
    df = pd.read_csv(SRC, index_col=0)
    out = {"source": "Python/alphas_dg.csv", "rows": int(len(df)), "groups": {}, "min_by_type_order": {}}
    for (typ, order), g in df.groupby(["type", "order"]):
        out["min_by_type_order"][f"{typ}:{int(order)}"] = float(g["alpha"].min())
    for (typ, ms, order), g in df.groupby(["type", "mesh_size", "order"]):
        out["groups"][f"{typ}:{ms}:{int(order)}"] = int(len(g))
    # right-isosceles P1 triangles: first rows of the coarsest ehdg mesh, order 1
    g = df[(df["type"] == "ehdg") & (df["mesh_size"] == 0.25) & (df["order"] == 1)]["alpha"].to_numpy()
    out["ehdg_0.25_p1_row0"] = float(g[0])
    out["ehdg_0.25_p1_row11"] = float(g[11])
    # unmarked elements: ehdg == hdg on the same mesh
    diffs = {}
    for ms in sorted(df["mesh_size"].unique()):
        for order in sorted(df["order"].unique()):
            a = df[(df["type"] == "ehdg") & (df["mesh_size"] == ms) & (df["order"] == order)]["alpha"].to_numpy()
            b = df[(df["type"] == "hdg") & (df["mesh_size"] == ms) & (df["order"] == order)]["alpha"].to_numpy()
            if len(a) == len(b) and len(a):
                diffs[f"{ms}:{int(order)}"] = int((abs(a - b) > 1e-9 * abs(b)).sum())
    out["n_elements_where_ehdg_differs_from_hdg"] = diffs
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
