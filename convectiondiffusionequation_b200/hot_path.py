"""Device pipeline of the (E)HDG hot path: thin Python over the C ABI (include/ehdg.h).

Stages mirror Convection_Diffusion._solveEHDG of the reference
(Python/convection_diffusion.py:223-403): mark -> prune -> alpha -> assemble+condense ->
CSR -> GMRES -> back-solve -> error.  torch is used for device buffers only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .coefficients import LayerExpr, SymExpr
from .enrichment_proxy import vertex_flags, vertex_flags_device
from .mesh import TriMesh


@dataclass
class ProblemSpec:
    order: int
    epsilon: float
    beta: tuple
    bonus_int: int = 10
    coeff: LayerExpr | None = None
    exact: LayerExpr | None = None
    enrich_functions: tuple = ()
    enrich_domain_ind: tuple = ()
    vol_drop_tol: float = 1e-6      # selection rule for dependent volume enrichments (ehdg_set_drop_tol); 0 = reference-literal

    def c_struct(self) -> _lib.Problem:
        p = _lib.Problem()
        p.order = int(self.order)
        p.bonus_int = int(self.bonus_int)
        p.epsilon = float(self.epsilon)
        p.beta[0], p.beta[1] = float(self.beta[0]), float(self.beta[1])
        if len(self.enrich_functions) > _lib.MAX_ENR:
            raise ValueError(f"at most {_lib.MAX_ENR} enrichment functions are supported")
        p.n_enrich = len(self.enrich_functions)
        for i, f in enumerate(self.enrich_functions):
            ax, k = f.single_axis()
            p.enrich_axis[i] = ax
            p.enrich_k[i] = k
        for name in ("coeff", "exact"):
            f = getattr(self, name)
            if f is not None and not isinstance(f, SymExpr):
                c, k = f.desc()
                for i in range(4):
                    getattr(p, name).c[i] = c[i]
                for i in range(2):
                    getattr(p, name).k[i] = k[i]
        return p


LAYER_DROP = 50.0      # EHDG_LAYER_DROP of csrc/ehdg_fast.h


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class EHDGPipeline:
    """One (mesh, order) instance of the hot path on one GPU."""

    def __init__(self, mesh: TriMesh, spec: ProblemSpec, device="cuda:0", force_general=False, dirichlet_lifting=False):
        if not torch.cuda.is_available():
            raise _lib.EhdgError("no CUDA device: the (E)HDG hot path has no CPU fallback")
        self.lib = _lib.load()
        self.dev = torch.device(device)
        torch.cuda.set_device(self.dev)
        self.mesh_host = mesh
        self.spec = spec
        self.force_general = bool(force_general)
        self.dirichlet_lifting = bool(dirichlet_lifting)    # boundary facets carry the projection of `exact` (debug_ehdg.py:210-214)
        self.ubnd = None
        self.nenr = len(spec.enrich_functions)
        self._cprob = spec.c_struct()
        self.ctx = C.c_void_p()
        _lib.check(self.lib.ehdg_create(C.byref(self._cprob), C.byref(self.ctx)), "ehdg_create")
        _lib.check(self.lib.ehdg_set_drop_tol(self.ctx, float(spec.vol_drop_tol)), "ehdg_set_drop_tol")
        # generic coefficient functions (run.py:29-30 as arbitrary expressions): postfix programs for the kernels
        progs = []
        if isinstance(spec.coeff, SymExpr):
            progs.append((0, spec.coeff))
        if isinstance(spec.exact, SymExpr):
            progs += [(1, spec.exact), (2, spec.exact.diff(0)), (3, spec.exact.diff(1))]
        for which, e in progs:
            ops, args = e.program()
            _lib.check(self.lib.ehdg_set_expression(self.ctx, which, ops.ctypes.data_as(C.c_void_p), args.ctypes.data_as(C.c_void_p),
                                                    len(ops)), "ehdg_set_expression")
        self.sizes = _lib.Sizes()
        _lib.check(self.lib.ehdg_get_sizes(self.ctx, C.byref(self.sizes)), "ehdg_get_sizes")
        self.plan = None
        self.csr = None
        self.upload_mesh()

    # ------------------------------------------------------------------ mesh
    def upload_mesh(self, pinned=None):
        m = self.mesh_host
        self.ubnd = None
        src = pinned or dict(verts=torch.from_numpy(m.verts), tris=torch.from_numpy(m.tris),
                             el2facet=torch.from_numpy(m.el2facet), facet2el=torch.from_numpy(m.facet2el))
        self.verts = src["verts"].to(self.dev, non_blocking=True)
        self.tris = src["tris"].to(self.dev, non_blocking=True)
        self.el2facet = src["el2facet"].to(self.dev, non_blocking=True)
        self.facet2el = src["facet2el"].to(self.dev, non_blocking=True)
        self.cmesh = _lib.Mesh(m.nv, m.ne, m.nf, self.verts.data_ptr(), self.tris.data_ptr(),
                               self.el2facet.data_ptr(), self.facet2el.data_ptr())
        self.h2d_bytes = sum(int(t.numel() * t.element_size()) for t in (self.verts, self.tris, self.el2facet, self.facet2el))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    # ------------------------------------------------------------------ a1
    def mark(self):
        m = self.mesh_host
        ne, nf = m.ne, m.nf
        self.elem_mask = torch.zeros(ne, dtype=torch.uint8, device=self.dev)
        self.facet_mask = torch.zeros((nf + 3) // 4 * 4, dtype=torch.uint8, device=self.dev)
        if self.nenr:
            self.vert_flags = vertex_flags_device(self.verts, self.spec.enrich_domain_ind, m.maxh)
            if self.vert_flags is None:            # predicate needs Python scalars: evaluate on the host copy
                vf = vertex_flags(m.verts, self.spec.enrich_domain_ind, m.maxh)
                self.vert_flags = torch.from_numpy(vf).to(self.dev)
            _lib.check(self.lib.ehdg_mark(self.ctx, C.byref(self.cmesh), _ptr(self.vert_flags), _ptr(self.elem_mask),
                                          _ptr(self.facet_mask), self._stream()), "ehdg_mark")
        return self.make_plan(None)

    def select_elements(self, f_lo: int, f_hi: int):
        """(elem_select, elem_owned) u8 [ne] for the facet range [f_lo, f_hi) (partitioned assembly)"""
        ne = self.mesh_host.ne
        sel = torch.empty(ne, dtype=torch.uint8, device=self.dev)
        own = torch.empty(ne, dtype=torch.uint8, device=self.dev)
        _lib.check(self.lib.ehdg_select_elements(self.ctx, C.byref(self.cmesh), int(f_lo), int(f_hi), _ptr(sel), _ptr(own),
                                                 self._stream()), "ehdg_select_elements")
        return sel, own

    def make_plan(self, elem_select=None):
        """(re)builds the plan; elem_select (u8 [ne] device tensor or None): partitioned assembly keeps only these elements"""
        if self.plan is not None:
            self.lib.ehdg_plan_destroy(self.plan)
        self.plan = C.c_void_p()
        self.elem_select = elem_select
        _lib.check(self.lib.ehdg_plan_create_select(self.ctx, C.byref(self.cmesh), _ptr(self.facet_mask), int(self.force_general),
                                                    _ptr(elem_select), C.byref(self.plan), self._stream()), "ehdg_plan_create_select")
        self.plan_sizes = _lib.PlanSizes()
        _lib.check(self.lib.ehdg_plan_get_sizes(self.plan, C.byref(self.plan_sizes)), "ehdg_plan_get_sizes")
        return self

    def eslot(self) -> torch.Tensor:
        """i32 [ne]: bit 30 = general path, low bits = index on that path (copy of the plan's array)."""
        torch.cuda.synchronize(self.dev)
        return _device_view(self.lib.ehdg_plan_eslot(self.plan), self.mesh_host.ne, torch.int32, self.dev).clone()

    # ------------------------------------------------------------------ a3
    def prune(self, threshold=1e-5, want_factors=False):
        m = self.mesh_host
        self.elem_active = torch.zeros(m.ne, dtype=torch.uint8, device=self.dev)
        self.facet_active = torch.zeros((m.nf + 3) // 4 * 4, dtype=torch.uint8, device=self.dev)
        self.factors = (torch.empty((m.ne, 4 * self.nenr), dtype=torch.float64, device=self.dev)
                        if (want_factors and self.nenr) else None)
        _lib.check(self.lib.ehdg_prune(self.ctx, self.plan, C.byref(self.cmesh), _ptr(self.elem_mask), _ptr(self.facet_mask),
                                       float(threshold), _ptr(self.elem_active), _ptr(self.facet_active), _ptr(self.factors),
                                       self._stream()), "ehdg_prune")
        return self

    # ------------------------------------------------------------------ a4
    def alpha(self):
        ne = self.mesh_host.ne
        if not hasattr(self, "lam") or self.lam.numel() != ne:
            self.lam = torch.empty(ne, dtype=torch.float64, device=self.dev)
            self.status = torch.zeros(ne, dtype=torch.int32, device=self.dev)
        _lib.check(self.lib.ehdg_alpha(self.ctx, self.plan, C.byref(self.cmesh), _ptr(self.elem_active), _ptr(self.lam),
                                       _ptr(self.status), self._stream()), "ehdg_alpha")
        return self

    # ------------------------------------------------------------------ a5-a7
    def alloc_blocks(self, want_full=False):
        ps = self.plan_sizes
        f64 = dict(dtype=torch.float64, device=self.dev)
        self.S = torch.empty(ps.S_doubles, **f64)
        self.g = torch.empty(ps.g_doubles, **f64)
        self.W = torch.empty(ps.W_doubles, **f64)
        self.z = torch.empty(ps.z_doubles, **f64)
        s = self.sizes
        self.afull_stride = s.nvmax * (s.nvmax + s.nfmax + 1) + s.nfmax * s.nvmax + s.nfmax * s.nfmax
        self.Afull = torch.zeros(ps.n_gen * self.afull_stride, **f64) if want_full else None

    def assemble_condense(self, want_full=False):
        if getattr(self, "S", None) is None or self.S.numel() != self.plan_sizes.S_doubles or (want_full and self.Afull is None):
            self.alloc_blocks(want_full)
        self.fused = False
        _lib.check(self.lib.ehdg_assemble_condense(self.ctx, self.plan, C.byref(self.cmesh), _ptr(self.lam), _ptr(self.elem_active),
                                                   _ptr(self.facet_mask), _ptr(self.S), _ptr(self.g), _ptr(self.W), _ptr(self.z),
                                                   _ptr(self.status), _ptr(self.Afull) if want_full else None, self._stream()),
                   "ehdg_assemble_condense")
        return self

    # ------------------------------------------------------------------ a8
    def csr_symbolic(self, nranks=1, rank=0):
        """pattern of the condensed system; nranks > 1: only the rows of this rank's facet range (partitioned assembly)"""
        if self.csr is not None:
            self.lib.ehdg_csr_destroy(self.csr)
        self.csr = C.c_void_p()
        self._lines_csr = None
        _lib.check(self.lib.ehdg_csr_symbolic_part(self.ctx, C.byref(self.cmesh), _ptr(self.facet_active) if self.nenr else None,
                                                   _ptr(self.facet_mask) if self.nenr else None, int(nranks), int(rank),
                                                   C.byref(self.csr), self._stream()), "ehdg_csr_symbolic_part")
        self.csr_sizes = _lib.CsrSizes()
        _lib.check(self.lib.ehdg_csr_get_sizes(self.csr, C.byref(self.csr_sizes)), "ehdg_csr_get_sizes")
        rb = np.zeros(nranks + 1, dtype=np.int64)
        fb = np.zeros(nranks + 1, dtype=np.int32)
        _lib.check(self.lib.ehdg_csr_get_partition(self.csr, rb.ctypes.data_as(C.c_void_p), fb.ctypes.data_as(C.c_void_p)),
                   "ehdg_csr_get_partition")
        self.row_begin, self.facet_begin = rb, fb
        lo, hi = C.c_int64(0), C.c_int64(0)
        _lib.check(self.lib.ehdg_csr_get_extent(self.csr, C.byref(lo), C.byref(hi)), "ehdg_csr_get_extent")
        self.own_extent = (int(lo.value), int(hi.value))    # columns the own rows reference (SpMV halo)
        return self

    def csr_alloc_values(self):
        """value array + right-hand side of the current pattern; storage of the right size is kept (a rebuilt pattern of the same
        mesh family must not pay a multi-GB allocation per mesh)"""
        n, nnz = self.csr_sizes.n_rows, self.csr_sizes.nnz
        if getattr(self, "vals", None) is None or self.vals.numel() != max(nnz, 1):
            self.vals = None
            self.vals = torch.empty(max(nnz, 1), dtype=torch.float64, device=self.dev)
        if getattr(self, "rhs", None) is None or self.rhs.numel() != max(n, 1):
            self.rhs = torch.zeros(max(n, 1), dtype=torch.float64, device=self.dev)
        return self

    def csr_build(self):
        if self.csr is not None:
            self.lib.ehdg_csr_destroy(self.csr)
        self.csr = C.c_void_p()
        self._lines_csr = None
        _lib.check(self.lib.ehdg_csr_symbolic(self.ctx, C.byref(self.cmesh), _ptr(self.facet_active) if self.nenr else None,
                                              _ptr(self.facet_mask) if self.nenr else None, C.byref(self.csr), self._stream()), "ehdg_csr_symbolic")
        self.csr_sizes = _lib.CsrSizes()
        _lib.check(self.lib.ehdg_csr_get_sizes(self.csr, C.byref(self.csr_sizes)), "ehdg_csr_get_sizes")
        n, nnz = self.csr_sizes.n_rows, self.csr_sizes.nnz
        self.vals = torch.empty(max(nnz, 1), dtype=torch.float64, device=self.dev)
        self.rhs = torch.empty(max(n, 1), dtype=torch.float64, device=self.dev)
        self.csr_numeric()
        return self

    def csr_pattern(self):
        """pattern + value storage without the numeric fill (the fused a5-a8 store fills it: assemble_condense_csr)"""
        if self.csr is not None:
            self.lib.ehdg_csr_destroy(self.csr)
        self.csr = C.c_void_p()
        self._lines_csr = None
        _lib.check(self.lib.ehdg_csr_symbolic(self.ctx, C.byref(self.cmesh), _ptr(self.facet_active) if self.nenr else None,
                                              _ptr(self.facet_mask) if self.nenr else None, C.byref(self.csr), self._stream()), "ehdg_csr_symbolic")
        self.csr_sizes = _lib.CsrSizes()
        _lib.check(self.lib.ehdg_csr_get_sizes(self.csr, C.byref(self.csr_sizes)), "ehdg_csr_get_sizes")
        return self.csr_alloc_values()

    def assemble_condense_csr(self):
        """a5-a8 in one pass: the element kernels store the condensed blocks straight into the value array of the pattern
        (ehdg_assemble_condense_csr); no element-wise S_K store, no separate numeric fill.  Same values as
        assemble_condense() + csr_numeric(); not available with Dirichlet lifting (the lifted right-hand side needs S_K)."""
        if self.dirichlet_lifting:
            raise ValueError("the fused store has no lifting variant: use assemble_condense() + csr_numeric()")
        if self.csr is None:
            self.csr_pattern()
        if getattr(self, "vals", None) is None or self.vals.numel() < max(self.csr_sizes.nnz, 1):
            self.csr_alloc_values()
        ps = self.plan_sizes
        if getattr(self, "W", None) is None or self.W.numel() != ps.W_doubles:
            f64 = dict(dtype=torch.float64, device=self.dev)
            self.W = torch.empty(ps.W_doubles, **f64)
            self.z = torch.empty(ps.z_doubles, **f64)
        self.S = self.g = None
        self.fused = True
        self.vals_folded = False
        _lib.check(self.lib.ehdg_assemble_condense_csr(self.ctx, self.plan, self.csr, C.byref(self.cmesh), _ptr(self.lam),
                                                       _ptr(self.elem_active), _ptr(self.facet_mask), _ptr(self.vals), _ptr(self.rhs),
                                                       _ptr(self.W), _ptr(self.z), _ptr(self.status), self._stream()),
                   "ehdg_assemble_condense_csr")
        return self

    def assemble_csr(self):
        """alpha + fused assembly/condensation/fill: the whole a4-a8 step"""
        return self.alpha().assemble_condense_csr()

    def dirichlet_values(self):
        """ubnd f64[nf][p+1]: L2 projection of `exact` on the boundary facets, zero elsewhere"""
        self.ubnd = torch.empty(self.mesh_host.nf * self.sizes.nfl, dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.ehdg_dirichlet_project(self.ctx, C.byref(self.cmesh), _ptr(self.ubnd), self._stream()), "ehdg_dirichlet_project")
        return self.ubnd

    def csr_numeric(self):
        self.vals_folded = False
        if getattr(self, "fused", False) and getattr(self, "S", None) is None:
            return self.assemble_condense_csr()              # fused pipeline: the values come from the element kernels
        if self.dirichlet_lifting and self.ubnd is None:
            self.dirichlet_values()
        _lib.check(self.lib.ehdg_csr_numeric_bc(self.ctx, self.plan, self.csr, C.byref(self.cmesh), _ptr(self.S), _ptr(self.g),
                                                _ptr(self.ubnd) if self.dirichlet_lifting else None, _ptr(self.vals), _ptr(self.rhs),
                                                self._stream()), "ehdg_csr_numeric_bc")
        return self

    def csr_arrays(self):
        """(rowptr i64, colind i32, row_start i32, dofmask u16) as torch tensors (copies)."""
        n, nnz, nf = self.csr_sizes.n_rows, self.csr_sizes.nnz, self.mesh_host.nf
        rp = _device_view(self.lib.ehdg_csr_rowptr(self.csr), n + 1, torch.int64, self.dev).clone()
        ci = _device_view(self.lib.ehdg_csr_colind(self.csr), max(nnz, 1), torch.int32, self.dev)[:nnz].clone()
        rs = _device_view(self.lib.ehdg_csr_row_start(self.csr), nf + 1, torch.int32, self.dev).clone()
        dm = _device_view(self.lib.ehdg_csr_dofmask(self.csr), nf, torch.int16, self.dev).clone()
        return rp, ci, rs, dm

    # ------------------------------------------------------------------ a9
    def spmv(self, x, y=None):
        if getattr(self, "vals_folded", False):
            self.csr_numeric()                      # a facet-mode solve left A M^-1 in vals
        y = torch.empty_like(x) if y is None else y
        _lib.check(self.lib.ehdg_spmv(self.ctx, self.csr, _ptr(self.vals), _ptr(x), _ptr(y), self._stream()), "ehdg_spmv")
        return y

    def set_lines(self, facet_line=None, direction=(0.0, 0.0), width=None, offset=0.25):
        """chains of the line-block preconditioner: facet_line (i32 [nf] device tensor: any user partition of the facets into
        chains) or, by default, lines of width `width` (default: the mesh size) across the flow direction (ehdg_csr_auto_lines)"""
        if facet_line is not None:
            fl = facet_line.to(self.dev, torch.int32).contiguous()
            _lib.check(self.lib.ehdg_csr_set_lines(self.csr, _ptr(fl), self._stream()), "ehdg_csr_set_lines")
        else:
            w = float(self.mesh_host.maxh if width is None else width)
            _lib.check(self.lib.ehdg_csr_auto_lines(self.ctx, self.csr, C.byref(self.cmesh), float(direction[0]), float(direction[1]),
                                                    w, float(offset), self._stream()), "ehdg_csr_auto_lines")
        self._lines_csr = self.csr.value
        return self

    def lines(self) -> torch.Tensor:
        return _device_view(self.lib.ehdg_csr_lines(self.csr), self.mesh_host.nf, torch.int32, self.dev).clone()

    def krylov_workspace_bytes(self, restart, method="dqgmres", precond="line", jacobi_sweeps=1) -> int:
        info = _lib.GmresInfo()
        info.restart, info.jacobi_sweeps = int(restart), int(jacobi_sweeps)
        info.use_line_block = int(precond == "line")
        info.use_strip_block = int(self.nenr > 0)
        out = C.c_int64(0)
        _lib.check(self.lib.ehdg_krylov_workspace_bytes(self.ctx, self.csr, C.byref(info), {"gmres": 0, "bicgstab": 1, "dqgmres": 2}[method],
                                                        C.byref(out)), "ehdg_krylov_workspace_bytes")
        return int(out.value)

    def set_workspace(self, nbytes: int):
        """hands the library a caller-owned device buffer for the Krylov vectors (no allocation inside the solve while it fits)"""
        self.workspace = torch.empty(int(nbytes), dtype=torch.uint8, device=self.dev) if nbytes else None
        _lib.check(self.lib.ehdg_set_workspace(self.ctx, _ptr(self.workspace), int(nbytes)), "ehdg_set_workspace")
        return self

    def gmres(self, restart=30, max_iters=20000, rtol=1e-12, strip_block=True, jacobi_sweeps=1, method="gmres", precond="line",
              line_target_dofs=0, drop_tol=0.0, line_max_blocks=-1):
        """Krylov solve of the condensed system.  precond = 'line' (line-block Jacobi: every chain of facets solved exactly; the
        matrix values stay untouched) or 'facet' (facet block-Jacobi folded into the matrix values: they hold A M^-1 afterwards and
        are refilled by csr_numeric before the next use)."""
        if precond not in ("line", "facet"):
            raise ValueError("precond must be 'line' or 'facet'")
        n = self.csr_sizes.n_rows
        if getattr(self, "vals_folded", False):
            self.csr_numeric()                      # an earlier facet-mode solve left A M^-1 in vals
        if precond == "line" and getattr(self, "_lines_csr", None) != self.csr.value:
            self.set_lines()
        self.xhat = torch.zeros(max(n, 1), dtype=torch.float64, device=self.dev)
        info = _lib.GmresInfo()
        info.restart, info.max_iters, info.rtol = int(restart), int(max_iters), float(rtol)
        info.use_strip_block = int(bool(strip_block) and self.nenr > 0)
        info.jacobi_sweeps = int(jacobi_sweeps)
        info.use_line_block = int(precond == "line")
        info.line_target_dofs = int(line_target_dofs)
        info.drop_tol = float(drop_tol)
        info.line_max_blocks = int(line_max_blocks)
        fn = {"gmres": self.lib.ehdg_gmres, "bicgstab": self.lib.ehdg_bicgstab, "dqgmres": self.lib.ehdg_dqgmres}[method]
        self.vals_folded = precond == "facet"       # set before the call: an error inside may leave vals half folded
        _lib.check(fn(self.ctx, self.csr, _ptr(self.vals), _ptr(self.rhs), _ptr(self.xhat), C.byref(info), self._stream()),
                   "ehdg_" + method)
        self.gmres_info = info
        return self

    def solve_auto(self, rtol=1e-8, max_iters=8000, precond="line", drop_tol=0.0, strip_block=True, line_max_blocks=-1):
        """Solver portfolio of the facet system (measured on B200, tools/line_probe.py): DQGMRES(8) on the line-block
        preconditioner is the fastest where it converges (no restart, eight vectors: 4M elements in 1 520 iterations / 15 s) but
        erratic on some systems; GMRES(m) is robust once m covers the plateau that the residual sits on until the information
        has crossed the mesh, m = #lines / 10 clamped to [30, 200] (4M elements: GMRES(30) stalls, GMRES(150) needs 1 937
        iterations / 45 s).  Plain systems try the first and fall back to the second; enriched systems take the second directly.
        self.solve_attempts records what ran."""
        self.solve_attempts = []
        if precond == "line":
            if getattr(self, "_lines_csr", None) != self.csr.value:
                self.set_lines()
            chains = int(torch.unique(self.lines()).numel())
        else:
            chains = 300
        m_auto = int(min(200, max(30, chains // 10)))
        # enriched systems: DQGMRES is unreliable on them (config 2 does not converge with it, config 3 takes 2.5x as long as
        # GMRES), so they go straight to GMRES; plain systems try the cheap method first
        # Enriched systems: GMRES(m), and once more with twice the restart if that does not converge within 4 000 iterations
        # (converging enriched solves need 700-1 900; what made config 2 stall in the r02d / r02g bench lines was a cut strip
        # chain, see chain_setup, not the restart length)
        m_enr = m_auto
        plan = (("dqgmres", 8), ("gmres", m_auto)) if self.nenr == 0 else (("gmres", m_enr), ("gmres", min(240, 2 * m_enr)))
        for k, (method, restart) in enumerate(plan):
            cap = max_iters
            if method == "dqgmres" and len(plan) > 1:
                cap = min(max_iters, 2500)                # where DQGMRES(8) works it is done well before
            elif self.nenr and k == 0 and len(plan) > 1:
                cap = min(max_iters, 4000)                # converging enriched solves need 700-1 900 iterations
            self.gmres(restart=restart, max_iters=cap, rtol=rtol, method=method, precond=precond, drop_tol=drop_tol,
                       strip_block=strip_block, line_max_blocks=line_max_blocks)
            gi = self.gmres_info
            self.solve_attempts.append(dict(method=method, restart=restart, iters=int(gi.iters), converged=bool(gi.converged),
                                            rel_residual=float(gi.rel_residual), seconds=float(gi.total_seconds)))
            if gi.converged:
                break
        return self

    # ------------------------------------------------------------------ a10
    def backsolve(self):
        self.u = torch.empty(self.plan_sizes.z_doubles, dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.ehdg_backsolve_bc(self.ctx, self.plan, self.csr, C.byref(self.cmesh), _ptr(self.W), _ptr(self.z),
                                              _ptr(self.xhat), _ptr(self.ubnd) if self.dirichlet_lifting else None, _ptr(self.u),
                                              self._stream()), "ehdg_backsolve_bc")
        return self

    def errors(self):
        """(L2 error, H1-seminorm error) against spec.exact with the order 50+bonus rule."""
        e = self.error_sums().cpu().numpy()
        return float(np.sqrt(e[0])), float(np.sqrt(e[1]))

    def error_sums(self, elem_select=None):
        """f64[2] device tensor: squared L2 and H1-seminorm errors summed over the plan's elements (or elem_select of them)"""
        err2 = torch.zeros(2, dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.ehdg_l2_error_select(self.ctx, self.plan, C.byref(self.cmesh), _ptr(self.elem_active), _ptr(elem_select),
                                                 _ptr(self.u), _ptr(err2), self._stream()), "ehdg_l2_error_select")
        return err2

    def moment_rhs_fraction(self) -> float:
        """share of the polynomial-path elements on which the exponentials of both layer functions of `coeff` are negligible, so that
        fast_assemble_kernel takes f_V from the moment vectors (same test as csrc/ehdg_fast.h; used for flop accounting)"""
        f = self.spec.coeff
        if f is None or isinstance(f, SymExpr) or self.plan_sizes.n_poly == 0:
            return 0.0
        c, k = f.desc()
        dmin = (1.0 - self.verts[self.tris.long()]).amin(dim=1)          # [ne, 2]
        ok = torch.ones(dmin.shape[0], dtype=torch.bool, device=self.dev)
        if c[1] != 0.0 or c[3] != 0.0:
            ok &= dmin[:, 0] * k[0] > LAYER_DROP
        if c[2] != 0.0 or c[3] != 0.0:
            ok &= dmin[:, 1] * k[1] > LAYER_DROP
        slot = _device_view(self.lib.ehdg_plan_eslot(self.plan), self.mesh_host.ne, torch.int32, self.dev)
        poly = (slot >= 0) & ((slot & 0x40000000) == 0)
        return float((ok & poly).sum().item()) / float(self.plan_sizes.n_poly)

    def kernel_timing(self, enable=True):
        _lib.check(self.lib.ehdg_kernel_timing(self.ctx, int(enable)), "ehdg_kernel_timing")
        return self

    def kernel_times_ms(self):
        """(alpha kernel, assemble+condense kernel) durations of the most recent polynomial-path launches"""
        a, b = C.c_float(0), C.c_float(0)
        _lib.check(self.lib.ehdg_kernel_times_ms(self.ctx, C.byref(a), C.byref(b)), "ehdg_kernel_times_ms")
        return float(a.value), float(b.value)

    # ------------------------------------------------------------------ whole path
    def setup(self, prune_threshold=1e-5):
        return self.mark().prune(prune_threshold)

    def assemble(self):
        """the metric's unit of work: alpha + element blocks + static condensation of every element"""
        return self.alpha().assemble_condense()

    def solve(self, **gm):
        """CSR build + Krylov solve (method = 'gmres' | 'dqgmres' | 'bicgstab') + back-solve"""
        return self.csr_build().gmres(**gm).backsolve()

    def ndof(self) -> int:
        """fes.ndof of the reference's compound space [V, F, Qx, QFx, Qy, QFy] (conv_diff.py:246, :396)."""
        m, s = self.mesh_host, self.sizes
        n = m.ne * s.n_p + m.nf * s.nfl
        if self.nenr:
            em = self.elem_mask.cpu().numpy()
            fm = self.facet_mask.cpu().numpy()[:m.nf]
            for i in range(self.nenr):
                n += int(((em >> i) & 1).sum()) + int(((fm >> i) & 1).sum())
        return int(n)

    def close(self):
        if self.csr is not None:
            self.lib.ehdg_csr_destroy(self.csr)
            self.csr = None
        if self.plan is not None:
            self.lib.ehdg_plan_destroy(self.plan)
            self.plan = None
        if self.ctx:
            self.lib.ehdg_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _device_view(ptr, n, dtype, dev):
    """torch view of a library-owned device buffer (no ownership transfer)."""
    itemsize = torch.empty((), dtype=dtype).element_size()

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {
        "shape": (int(n),), "typestr": {torch.int32: "<i4", torch.int64: "<i8", torch.int16: "<i2",
                                        torch.float64: "<f8", torch.uint8: "|u1"}[dtype],
        "data": (int(ptr), False), "version": 2, "strides": (itemsize,)}
    return torch.as_tensor(h, device=dev)
