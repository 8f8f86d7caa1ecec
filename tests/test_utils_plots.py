"""Reporting parity (reference Python/utils.py): the two error plots exist, headless, with one series per Type and panel."""
import pandas as pd

from convectiondiffusionequation_b200.utils import plot_error_dof, plot_error_mesh


def _table():
    rows = []
    for order in (1, 2, 3, 4):
        for typ, c in (("hdg", 1.0), ("ehdg", 0.3)):
            for h in (0.25, 0.125, 0.0625):
                rows.append([order, int(100 * order / h ** 2), h, c * h ** (order + 1), 10, typ])
    return pd.DataFrame(rows, columns=['Order', 'DOFs', 'Mesh Size', 'Error', 'Bonus Int', 'Type'])


def test_plots_are_written(tmp_path):
    t = _table()
    for fn, name in ((plot_error_dof, "dof"), (plot_error_mesh, "mesh")):
        out = fn(t, str(tmp_path / name))
        data = open(out, "rb").read()
        assert len(data) > 1000
        if out.endswith(".svg"):
            txt = data.decode()
            assert txt.count('class="series"') == 8           # 4 panels x 2 types
            assert 'p = 4' in txt and 'ehdg' in txt
