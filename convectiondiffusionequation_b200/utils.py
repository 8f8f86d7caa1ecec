"""Error plots of the reference's Python/utils.py:4-49: four log-log panels (p = 1..4), one '.-' series per 'Type', against the
DOFs (`plot_error_dof`) or the mesh size (`plot_error_mesh`).  Headless: with matplotlib the figure is built as in the reference and
saved instead of shown; without it (this image) the same picture is written as a self-contained SVG.  Both return the file path."""
from __future__ import annotations

import math

PANEL_W, PANEL_H, MARGIN_L, MARGIN_B, MARGIN_T, GAP = 300, 260, 62, 44, 30, 24
COLOURS = ["#1f77b4", "#ff7f0e", "#2ca02c", "#d62728", "#9467bd", "#8c564b"]


def _series(table, order, xcol):
    sub = table[table.Order.eq(float(order))]
    out = []
    for name, group in sub.groupby('Type'):
        g = group.sort_values(xcol)
        pts = [(float(a), float(b)) for a, b in zip(g[xcol], g['Error']) if float(a) > 0 and float(b) > 0 and math.isfinite(float(b))]
        if pts:
            out.append((str(name), pts))
    return out


def _decades(lo, hi):
    a, b = math.floor(math.log10(lo)), math.ceil(math.log10(hi))
    if a == b:
        a, b = a - 1, b + 1
    return a, b


def _svg(table, xcol, path):
    W = MARGIN_L + 4 * (PANEL_W + GAP)
    H = MARGIN_T + PANEL_H + MARGIN_B
    el = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{W}" height="{H}" font-family="sans-serif" font-size="11">',
          f'<rect width="{W}" height="{H}" fill="white"/>']
    for pi, order in enumerate((1, 2, 3, 4)):
        x0 = MARGIN_L + pi * (PANEL_W + GAP)
        y0 = MARGIN_T
        ser = _series(table, order, xcol)
        el.append(f'<rect x="{x0}" y="{y0}" width="{PANEL_W}" height="{PANEL_H}" fill="none" stroke="black"/>')
        el.append(f'<text x="{x0 + PANEL_W / 2}" y="{y0 - 8}" text-anchor="middle" font-size="13">p = {order}</text>')
        el.append(f'<text x="{x0 + PANEL_W / 2}" y="{y0 + PANEL_H + 34}" text-anchor="middle">{xcol}</text>')
        if pi == 0:
            el.append(f'<text x="14" y="{y0 + PANEL_H / 2}" text-anchor="middle" transform="rotate(-90 14 {y0 + PANEL_H / 2})">Error</text>')
        if not ser:
            continue
        xs = [p[0] for _, pts in ser for p in pts]
        ys = [p[1] for _, pts in ser for p in pts]
        xa, xb = _decades(min(xs), max(xs))
        ya, yb = _decades(min(ys), max(ys))
        fx = lambda v: x0 + (math.log10(v) - xa) / (xb - xa) * PANEL_W
        fy = lambda v: y0 + PANEL_H - (math.log10(v) - ya) / (yb - ya) * PANEL_H
        for d in range(xa, xb + 1):
            X = x0 + (d - xa) / (xb - xa) * PANEL_W
            el.append(f'<line x1="{X:.1f}" y1="{y0}" x2="{X:.1f}" y2="{y0 + PANEL_H}" stroke="#dddddd"/>')
            el.append(f'<text x="{X:.1f}" y="{y0 + PANEL_H + 14}" text-anchor="middle">1e{d}</text>')
        for d in range(ya, yb + 1):
            Y = y0 + PANEL_H - (d - ya) / (yb - ya) * PANEL_H
            el.append(f'<line x1="{x0}" y1="{Y:.1f}" x2="{x0 + PANEL_W}" y2="{Y:.1f}" stroke="#dddddd"/>')
            el.append(f'<text x="{x0 - 4}" y="{Y + 4:.1f}" text-anchor="end">1e{d}</text>')
        for si, (name, pts) in enumerate(ser):
            col = COLOURS[si % len(COLOURS)]
            path_d = " ".join(f"{'M' if i == 0 else 'L'}{fx(a):.1f},{fy(b):.1f}" for i, (a, b) in enumerate(pts))
            el.append(f'<path d="{path_d}" fill="none" stroke="{col}" stroke-width="1.5" class="series" data-type="{name}"/>')
            for a, b in pts:
                el.append(f'<circle cx="{fx(a):.1f}" cy="{fy(b):.1f}" r="2.5" fill="{col}"/>')
            ly = y0 + 14 + 14 * si
            el.append(f'<line x1="{x0 + PANEL_W - 70}" y1="{ly - 4}" x2="{x0 + PANEL_W - 50}" y2="{ly - 4}" stroke="{col}" stroke-width="1.5"/>')
            el.append(f'<text x="{x0 + PANEL_W - 46}" y="{ly}">{name}</text>')
    el.append('</svg>')
    with open(path, "w") as f:
        f.write("\n".join(el))
    return path


def _plot(table, xcol, path):
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        return _svg(table, xcol, path if path.endswith(".svg") else path + ".svg")
    fig, axes = plt.subplots(1, 4, figsize=(20, 5))
    axes[0].set_ylabel("Error")
    for ax, order in zip(axes, (1, 2, 3, 4)):
        for name, group in table[table.Order.eq(float(order))].groupby('Type'):
            group.plot(x=xcol, y='Error', ax=ax, label=name, legend=True, title=f'p = {order}', style='.-', loglog=True)
    out = path if path.rsplit(".", 1)[-1] in ("png", "pdf", "svg") else path + ".png"
    fig.savefig(out)
    plt.close(fig)
    return out


def plot_error_dof(table, path="error_dof"):
    """utils.py:4-25"""
    return _plot(table, 'DOFs', path)


def plot_error_mesh(table, path="error_mesh"):
    """utils.py:28-49"""
    return _plot(table, 'Mesh Size', path)
