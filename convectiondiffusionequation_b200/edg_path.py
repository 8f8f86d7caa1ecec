"""Device pipeline of the EDG path (enriched interior-penalty DG): thin Python over the C ABI (include/ehdg.h).

Stages mirror Convection_Diffusion._solveEDG of the reference (Python/convection_diffusion.py:42-219):
mark (:52-54) -> prune (:83-123) -> alpha (:129-163) -> pattern + forms + assembly (:165-198) -> solve (:202-203, PARDISO there,
line-block-Jacobi GMRES here) -> recombination + error (:205-210).  torch is used for device buffers only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .hot_path import EHDGPipeline, _ptr


class EDGPipeline(EHDGPipeline):
    """One (mesh, order) instance of the EDG path on one GPU.  Marking, the plan and the Krylov entry points are those of the
    EHDG pipeline (the matrix container is shared); pruning, alpha rule, pattern, assembly and error are EDG's own."""

    def __init__(self, *args, boundary_data=False, **kw):
        """boundary_data: weakly imposed inhomogeneous Dirichlet data on the right-hand side (reference debug_edg.py:191-192)"""
        super().__init__(*args, **kw)
        self.boundary_data = bool(boundary_data or self.dirichlet_lifting)
        self.dirichlet_lifting = False           # the facet lifting is an EHDG notion; EDG imposes the data weakly
        _lib.check(self.lib.ehdg_edg_set_boundary_data(self.ctx, int(self.boundary_data)), "ehdg_edg_set_boundary_data")

    def prune(self, theta=1e-5, want_factors=False):
        m = self.mesh_host
        self.elem_active = torch.zeros(m.ne, dtype=torch.uint8, device=self.dev)
        self.factors = (torch.empty((m.ne, self.nenr), dtype=torch.float64, device=self.dev) if (want_factors and self.nenr) else None)
        _lib.check(self.lib.ehdg_edg_prune(self.ctx, self.plan, C.byref(self.cmesh), _ptr(self.elem_mask), float(theta),
                                           _ptr(self.elem_active), _ptr(self.factors), self._stream()), "ehdg_edg_prune")
        return self

    def setup(self, theta=1e-5):
        return self.mark().prune(theta)

    def alpha(self):
        ne = self.mesh_host.ne
        if not hasattr(self, "lam") or self.lam.numel() != ne:
            self.lam = torch.empty(ne, dtype=torch.float64, device=self.dev)
            self.status = torch.zeros(ne, dtype=torch.int32, device=self.dev)
        _lib.check(self.lib.ehdg_edg_alpha(self.ctx, self.plan, C.byref(self.cmesh), _ptr(self.elem_active), _ptr(self.lam),
                                           _ptr(self.status), self._stream()), "ehdg_edg_alpha")
        return self

    def csr_symbolic(self, nranks=1, rank=0):
        if nranks != 1:
            raise _lib.EhdgError("the EDG path runs on one GPU")
        if self.csr is not None:
            self.lib.ehdg_csr_destroy(self.csr)
        self.csr = C.c_void_p()
        self._lines_csr = None
        _lib.check(self.lib.ehdg_edg_csr_symbolic(self.ctx, C.byref(self.cmesh), _ptr(self.elem_active) if self.nenr else None,
                                                  C.byref(self.csr), self._stream()), "ehdg_edg_csr_symbolic")
        self.csr_sizes = _lib.CsrSizes()
        _lib.check(self.lib.ehdg_csr_get_sizes(self.csr, C.byref(self.csr_sizes)), "ehdg_csr_get_sizes")
        return self

    def assemble(self):
        """alpha + pattern + values of every block row + right-hand side"""
        self.alpha()
        if self.csr is None:
            self.csr_symbolic()
        n, nnz = self.csr_sizes.n_rows, self.csr_sizes.nnz
        if not hasattr(self, "vals") or self.vals.numel() != max(nnz, 1):
            self.vals = torch.empty(max(nnz, 1), dtype=torch.float64, device=self.dev)
            self.rhs = torch.empty(max(n, 1), dtype=torch.float64, device=self.dev)
        return self.csr_numeric()

    def csr_numeric(self):
        self.vals_folded = False
        _lib.check(self.lib.ehdg_edg_assemble(self.ctx, C.byref(self.cmesh), self.csr, _ptr(self.lam), _ptr(self.elem_active),
                                              _ptr(self.vals), _ptr(self.rhs), _ptr(self.status), self._stream()), "ehdg_edg_assemble")
        return self

    def csr_build(self):
        return self.assemble()

    def assemble_condense(self, want_full=False):
        raise _lib.EhdgError("EDG has no static condensation: use assemble()")

    def gmres(self, **kw):
        kw.setdefault("precond", "line")
        if kw["precond"] != "line":
            raise ValueError("the EDG system is solved with the line-block preconditioner")
        kw["strip_block"] = False
        return super().gmres(**kw)

    def solve_auto(self, rtol=1e-8, max_iters=8000, precond="line", drop_tol=0.0, strip_block=False, line_max_blocks=-1):
        """GMRES on lines of elements.  Preconditioned by exact solves of the streamline lines, the iteration count grows with the
        number of lines (host study tests/studies/edg_gs_study.py: 25 / 46 / 127 iterations for 8 / 16 / 32 lines; a Gauss-Seidel
        sweep across the lines still needs about #lines): across the lines the SIP penalty eps 4 alpha p^2 / h^2 couples like a
        1-D elliptic operator.  The restart length therefore covers the lines, m = #lines + 20, at most 300; beyond a few dozen
        lines the EDG solve needs a coarse level -- N = 224 does not converge."""
        if getattr(self, "_lines_csr", None) != self.csr.value:
            self.set_lines()
        chains = int(torch.unique(self.lines()).numel())
        m = int(min(300, max(30, chains + 20)))
        self.gmres(restart=m, max_iters=max_iters, rtol=rtol, method="gmres", drop_tol=drop_tol, line_max_blocks=0)
        gi = self.gmres_info
        self.solve_attempts = [dict(method="gmres", restart=m, iters=int(gi.iters), converged=bool(gi.converged),
                                    rel_residual=float(gi.rel_residual), seconds=float(gi.total_seconds))]
        return self

    def csr_arrays(self):
        """(rowptr i64, colind i32, row_start i32 [ne+1]) as torch tensors (copies)"""
        from .hot_path import _device_view
        n, nnz, ne = self.csr_sizes.n_rows, self.csr_sizes.nnz, self.mesh_host.ne
        rp = _device_view(self.lib.ehdg_csr_rowptr(self.csr), n + 1, torch.int64, self.dev).clone()
        ci = _device_view(self.lib.ehdg_csr_colind(self.csr), max(nnz, 1), torch.int32, self.dev)[:nnz].clone()
        rs = _device_view(self.lib.ehdg_csr_row_start(self.csr), ne + 1, torch.int32, self.dev).clone()
        return rp, ci, rs

    def lines(self):
        """chain id of every ELEMENT (the blocks of the EDG matrix are elements)"""
        from .hot_path import _device_view
        return _device_view(self.lib.ehdg_csr_lines(self.csr), self.mesh_host.ne, torch.int32, self.dev).clone()

    def backsolve(self):
        """u [ne][n_p + n_enrich] from the solution vector (the recombination of :205-206 happens in the error kernel)"""
        s = self.sizes
        self.u = torch.empty(self.mesh_host.ne * s.nvmax, dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.ehdg_edg_unpack(self.ctx, C.byref(self.cmesh), self.csr, _ptr(self.elem_active), _ptr(self.xhat), _ptr(self.u),
                                            self._stream()), "ehdg_edg_unpack")
        return self

    def error_sums(self, elem_select=None):
        err2 = torch.zeros(2, dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.ehdg_edg_l2_error(self.ctx, C.byref(self.cmesh), _ptr(self.elem_active), _ptr(self.u), _ptr(err2), self._stream()),
                   "ehdg_edg_l2_error")
        return err2

    def solve(self, **gm):
        return self.assemble().gmres(**gm).backsolve()

    def ndof(self) -> int:
        """fes.ndof of the compound space [V, Q_0, Q_1, ...] (:56, :214)"""
        m, s = self.mesh_host, self.sizes
        n = m.ne * s.n_p
        if self.nenr:
            em = self.elem_mask.cpu().numpy()
            for i in range(self.nenr):
                n += int(((em >> i) & 1).sum())
        return int(n)
