/* libehdg -- C ABI of the B200-native (E)HDG convection-diffusion hot path.
 *
 * The reference (aqibrahimbt/convectiondiffusionequation) has no FFI of its own: its hot path
 * is a Python method that scripts NGSolve.  Each entry point below replaces one stage of
 *   Python/convection_diffusion.py:223-403  (Convection_Diffusion._solveEHDG)
 * and cites the lines it stands for.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every function returns 0 on success, a negative ehdg_status otherwise; ehdg_last_error()
 *     gives the text.  No exceptions cross the boundary.
 *   - all array pointers are DEVICE pointers owned by the caller unless the name says `host`;
 *     `stream` is a cudaStream_t passed as void* (0 = default stream).  Work is stream ordered.
 *   - the library owns only: the context (quadrature rules and reference tensors on the
 *     device), the plan (element classification lists), and scratch it allocates per call.
 *   - there is no CPU code path: without a CUDA device ehdg_create fails with EHDG_ERR_CUDA.
 */
#ifndef EHDG_H
#define EHDG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EHDG_MAX_ENR 2

typedef enum {
    EHDG_OK = 0,
    EHDG_ERR_ARG = -1,
    EHDG_ERR_CUDA = -2,
    EHDG_ERR_ALLOC = -3,
    EHDG_ERR_SINGULAR = -4,
    EHDG_ERR_NOCONV = -5,
    EHDG_ERR_NCCL = -6
} ehdg_status;

/* f(x,y) = c0 + c1 L(k0;x) + c2 L(k1;y) + c3 L(k0;x) L(k1;y),
 * L(k;s) = s + (exp(k(s-1)) - exp(-k)) / (exp(-k) - 1)   (run.py:26-27, k = beta_i/eps) */
typedef struct {
    double c[4];
    double k[2];
} ehdg_layer_expr;

/* Problem definition: the config keys of run.py:32-46 that reach the kernels. */
typedef struct {
    int32_t order;                   /* config['order'] entry, 1..6 */
    int32_t bonus_int;               /* config['bonus_int'] */
    double epsilon;                  /* config['epsilon'] */
    double beta[2];                  /* config['beta'] */
    int32_t n_enrich;                /* len(config['enrich_functions']), 0..2 */
    int32_t enrich_axis[EHDG_MAX_ENR]; /* 0: psi = L(k; x), 1: psi = L(k; y) */
    double enrich_k[EHDG_MAX_ENR];
    ehdg_layer_expr coeff;           /* config['coeff']  (right-hand side) */
    ehdg_layer_expr exact;           /* config['exact']  (error computation) */
} ehdg_problem;

/* Triangle mesh with facet topology (replaces Mesh(unit_square.GenerateMesh(maxh)),
 * convection_diffusion.py:228).  Local edges follow NGSolve ET_TRIG: e0={V2,V0}, e1={V1,V2},
 * e2={V0,V1}. */
typedef struct {
    int32_t nv, ne, nf;
    const double* verts;       /* f64 [nv,2] */
    const int32_t* tris;       /* i32 [ne,3] */
    const int32_t* el2facet;   /* i32 [ne,3] facet of local edge 0,1,2 */
    const int32_t* facet2el;   /* i32 [nf,2] adjacent elements ascending, -1 = none (boundary) */
} ehdg_mesh;

typedef struct ehdg_ctx ehdg_ctx;

/* derived sizes, filled by ehdg_get_sizes */
typedef struct {
    int32_t n_p;      /* (p+1)(p+2)/2 polynomial volume dofs */
    int32_t nfl;      /* p+1 polynomial dofs per facet */
    int32_t nb;       /* nfl + n_enrich dof slots per facet */
    int32_t nvmax;    /* n_p + n_enrich volume slots of a general element */
    int32_t nfmax;    /* 3 nb facet slots of a general element */
    int32_t nf_poly;  /* 3 nfl facet dofs of a polynomial (fast path) element */
    int32_t nq, nqf;  /* points of the order 2p+bonus triangle / segment rules */
    int32_t nq0, nqf0;/* same for order 2p (alpha integrators, convection_diffusion.py:315-320) */
    int32_t nq_err;   /* points of the order 50+bonus rule (convection_diffusion.py:393-394) */
    int64_t flops_poly_elem;   /* flops the polynomial path spends per element (alpha + assemble + condense) */
    int64_t flops_quad_elem;   /* SURVEY 8(d) quadrature-algorithm flops per unenriched element */
    int64_t flops_poly_alpha;    /* ... of which the alpha kernel */
    int64_t flops_poly_assemble; /* ... and the assemble + condense kernel */
} ehdg_sizes;

const char* ehdg_version(void);
const char* ehdg_last_error(void);

/* context: rules + reference tensors for (order, bonus, enrichment) on the current device.  A context owns small device-side
 * scratch (work counters of the element kernels, timing events, the grow-only plan scratch): calls on ONE context must be issued
 * from one host thread into one stream at a time; use one context per concurrently running stream. */
int ehdg_create(const ehdg_problem* prob, ehdg_ctx** out);
int ehdg_destroy(ehdg_ctx* ctx);
int ehdg_get_sizes(const ehdg_ctx* ctx, ehdg_sizes* out);
/* Generic coefficient functions (reference run.py:26-30, :44: `coeff` and `exact` are arbitrary NGSolve CoefficientFunctions
 * of x, y).  An expression is handed over as a postfix program of at most 128 instructions, ops[i] in
 *   1 CONST(args[i]) 2 X 3 Y 4 ADD 5 SUB 6 MUL 7 DIV 8 NEG 9 EXP 10 LOG 11 SIN 12 COS 13 TAN 14 SQRT 15 POWI(int args[i]) 16 POW
 *   17 TANH 18 ABS 19 ATAN 20 SIGN
 * and every kernel on the path evaluates it with a register stack (csrc/ehdg_expr.h).  which: 0 = coeff, 1 = exact, 2 / 3 =
 * d exact / dx, dy (H1 error).  n = 0 restores the built-in boundary-layer family of ehdg_problem.  The enrichment functions
 * stay in that family. */
int ehdg_set_expression(ehdg_ctx* ctx, int which, const int32_t* ops, const double* args, int n);
/* Selection rule for dependent VOLUME enrichment dofs inside the element condensation (a6): after the polynomial dofs of an
 * element have been eliminated, an enrichment whose remaining diagonal is below vol_drop_tol x its own size is dependent on
 * that element (e.g. q(y) where beta_y h / eps < 1) and leaves the element space: coefficient 0, no contribution to S_K;
 * status bit 8 marks the element.  Default 1e-6; 0 restores the reference's literal behaviour (every unpruned enrichment is
 * kept, A_VV may be singular to working precision).  The facet-level counterpart is ehdg_gmres_info.drop_tol. */
int ehdg_set_drop_tol(ehdg_ctx* ctx, double vol_drop_tol);
/* CUDA-event timing of the two polynomial-path kernels (recorded on the launching stream); off by default.
 * ehdg_kernel_times_ms returns the durations of the most recent launches. */
int ehdg_kernel_timing(ehdg_ctx* ctx, int enable);
int ehdg_kernel_times_ms(ehdg_ctx* ctx, float* alpha_ms, float* assemble_ms);

/* a1  mark_elements / mark_element_bnd  (enrichment_proxy.py:201-208, :228-251)
 * vert_flags u8[n_enrich][nv]: the Python predicates evaluated on the vertices (host side).
 * elem_mask u8[ne]: bit k = element marked for enrichment k (any vertex flagged).
 * facet_mask u8[nf]: bit k = facet belongs to an element marked for k. */
int ehdg_mark(ehdg_ctx* ctx, const ehdg_mesh* mesh, const uint8_t* vert_flags, uint8_t* elem_mask,
              uint8_t* facet_mask, void* stream);

/* plan: splits the elements into the polynomial path (no enriched dof anywhere on the element)
 * and the general quadrature path (marked elements and their facet neighbours) and assigns
 * each an output slot.  force_general != 0 sends every element through the quadrature path
 * (cross-checks, order sweeps of the prescribed quadrature algorithm). */
typedef struct ehdg_plan ehdg_plan;
typedef struct {
    int64_t n_poly, n_gen;        /* elements on each path */
    int64_t S_doubles, g_doubles; /* sizes of the condensed-block stores */
    int64_t W_doubles, z_doubles; /* sizes of the back-solve stores (= size of the solution u) */
} ehdg_plan_sizes;
int ehdg_plan_create(ehdg_ctx* ctx, const ehdg_mesh* mesh, const uint8_t* facet_mask, int force_general,
                     ehdg_plan** out, void* stream);
/* partitioned assembly: only the elements with elem_select[el] != 0 (u8 [ne], device; null = all) get a slot;
 * every later stage (alpha, assemble + condense, CSR numeric, back-solve, error) then works on that subset and the
 * element stores are sized by it.  A rank selects the elements that touch its facet range, i.e. its own elements plus
 * one layer of redundantly computed ghosts, so that no S_K crosses GPUs. */
/* elem_select[el] = 1 iff element el touches a facet of [f_lo, f_hi); elem_owned[el] (may be null) = 1 iff its
 * lowest-numbered facet lies in the range (every element has exactly one owner, and the owner has it selected) */
int ehdg_select_elements(ehdg_ctx* ctx, const ehdg_mesh* mesh, int32_t f_lo, int32_t f_hi, uint8_t* elem_select,
                         uint8_t* elem_owned, void* stream);
int ehdg_plan_create_select(ehdg_ctx* ctx, const ehdg_mesh* mesh, const uint8_t* facet_mask, int force_general,
                            const uint8_t* elem_select, ehdg_plan** out, void* stream);
int ehdg_plan_destroy(ehdg_plan* plan);
int ehdg_plan_get_sizes(const ehdg_plan* plan, ehdg_plan_sizes* out);
/* eslot i32[ne]: bit 30 set = general path; low bits = index on that path (device pointer, plan-owned) */
const int32_t* ehdg_plan_eslot(const ehdg_plan* plan);

/* a3  enrichment pruning (convection_diffusion.py:270-310), threshold 1e-5 there.
 * In : elem_mask, facet_mask from ehdg_mark.
 * Out: elem_active u8[ne]  = elem_mask minus pruned volume enrichment dofs,
 *      facet_active u8[nf] = facet_mask minus pruned facet enrichment dofs minus boundary facets
 *                            (dirichlet=".*", :230-232); nf rounded up to a multiple of 4 bytes,
 *      factors f64[ne][4*n_enrich] (optional, may be null): the reference's `factor` per visited
 *      dof in order Q_k, QF_k(e0,e1,e2); -1 where the dof does not exist / element not visited,
 *      -3 where the enrichment function vanishes identically on the facet (skipped). */
int ehdg_prune(ehdg_ctx* ctx, const ehdg_plan* plan, const ehdg_mesh* mesh, const uint8_t* elem_mask,
               const uint8_t* facet_mask, double threshold, uint8_t* elem_active, uint8_t* facet_active,
               double* factors, void* stream);

/* a4  stabilisation eigenvalue lambda_K = max eig(pinv(a + f f^T) m)  (convection_diffusion.py:314-351);
 * alpha_K = 20 lambda_K is applied inside ehdg_assemble_condense (:349).
 * lam f64[ne], status i32[ne] (optional). */
int ehdg_alpha(ehdg_ctx* ctx, const ehdg_plan* plan, const ehdg_mesh* mesh, const uint8_t* elem_active,
               double* lam, int32_t* status, void* stream);

/* a5+a6+a7  element blocks, right-hand side and static condensation
 * (convection_diffusion.py:357-383; condensation = NGSolve BilinearForm(condense=True)).
 * S,g: condensed blocks  S_K = A_FF - A_FV A_VV^-1 A_VF,  g_K = -A_FV A_VV^-1 f_V
 * W,z: back-solve operators  A_VV^-1 A_VF,  A_VV^-1 f_V.
 * Layout: polynomial elements first ([n_poly][nf_poly^2] ...), then general elements
 * ([n_gen][nfmax^2] ...); an element's index on its path is eslot & 0x3fffffff.
 * Afull (optional, general path only, [n_gen][nvmax*(nvmax+nfmax+1) + nfmax*nvmax + nfmax*nfmax]):
 * uncondensed blocks [A_VV | A_VF | f_V], A_FV, A_FF for parity tests. */
int ehdg_assemble_condense(ehdg_ctx* ctx, const ehdg_plan* plan, const ehdg_mesh* mesh, const double* lam,
                           const uint8_t* elem_active, const uint8_t* facet_mask, double* S, double* g,
                           double* W, double* z, int32_t* status, double* Afull, void* stream);

/* a8  condensed facet system in CSR.  Rows/columns = active facet dofs, facet-major
 * (row of (facet f, slot k) = row_start[f] + rank of k among the active slots of f),
 * columns ascending in every row, structural zeros kept (NGSolve graph rule).
 * Polynomial dofs of interior facets are free; enriched dof j of facet f is active iff bit j of
 * facet_active[f] (from ehdg_prune; pass null when n_enrich = 0).  facet_mask (from ehdg_mark, may be null)
 * tells which facets belong to marked elements: they form the strip block of the preconditioner. */
typedef struct ehdg_csr ehdg_csr;
typedef struct {
    int64_t n_rows, nnz;
} ehdg_csr_sizes;
int ehdg_csr_symbolic(ehdg_ctx* ctx, const ehdg_mesh* mesh, const uint8_t* facet_active, const uint8_t* facet_mask,
                      ehdg_csr** out, void* stream);
/* Row-partitioned variant: the rows are split into nranks nearly equal contiguous ranges at facet boundaries (the same
 * partition on every rank: it follows from the replicated dof masks) and this rank keeps the pattern of its own rows only.
 * n_rows stays the global size and column indices stay global; nnz, colind and the value array cover the own rows; the
 * right-hand side and the solution are indexed by the global row.  ehdg_csr_get_partition: row_begin i64[nranks+1],
 * facet_begin i32[nranks+1] (host).  The Krylov solvers accept such a matrix once ehdg_comm_set_partition holds the
 * same row_begin; they exchange the block-Jacobi inverses of the halo facets once and the halo of x per SpMV. */
int ehdg_csr_symbolic_part(ehdg_ctx* ctx, const ehdg_mesh* mesh, const uint8_t* facet_active, const uint8_t* facet_mask,
                           int nranks, int rank, ehdg_csr** out, void* stream);
int ehdg_csr_get_partition(const ehdg_csr* csr, int64_t* row_begin, int32_t* facet_begin);
/* column range [col_lo, col_hi) the rows of this rank reference (its SpMV halo extent), computed with the pattern */
int ehdg_csr_get_extent(const ehdg_csr* csr, int64_t* col_lo, int64_t* col_hi);
/* Chains of the line-block Jacobi preconditioner.  facet_line i32[nf] (device): chain id of every facet, any non-negative
 * integers; a negative entry makes the facet a chain of its own (= facet block-Jacobi for it).  The array is copied.
 * ehdg_csr_auto_lines fills it from the geometry: chain of facet f = floor(n . c_f / width + offset) with c_f the facet
 * midpoint and n = (dir_x, dir_y) a direction ACROSS the flow; dir = (0,0) picks n from beta (snapped to a coordinate
 * axis when beta is within 1:10 of it, so that lattice rows stay whole); width <= 0: sqrt(2 |Omega| / ne).
 * These replace nothing in the reference (it calls a direct solver, convection_diffusion.py:385-387); they parameterise
 * the preconditioner of ehdg_gmres / ehdg_dqgmres / ehdg_bicgstab. */
int ehdg_csr_set_lines(ehdg_csr* csr, const int32_t* facet_line, void* stream);
int ehdg_csr_auto_lines(ehdg_ctx* ctx, ehdg_csr* csr, const ehdg_mesh* mesh, double dir_x, double dir_y, double width,
                        double offset, void* stream);
const int32_t* ehdg_csr_lines(const ehdg_csr* csr);      /* i32 [nf] device, null until set */
int ehdg_csr_destroy(ehdg_csr* csr);
int ehdg_csr_get_sizes(const ehdg_csr* csr, ehdg_csr_sizes* out);
const int64_t* ehdg_csr_rowptr(const ehdg_csr* csr);     /* i64 [n_rows+1] device */
/* i32 [nnz] device.  ehdg_csr_symbolic writes the column indices of the FIRST row of every facet only (all rows of a facet share
 * one column set and the facet-block SpMV reads just that row); the first call of this accessor -- or the first internal
 * consumer of whole rows: the chain factorisation, the plain CSR SpMV -- fills the rest on the pattern's stream and waits for it. */
const int32_t* ehdg_csr_colind(const ehdg_csr* csr);
const int32_t* ehdg_csr_row_start(const ehdg_csr* csr);  /* i32 [nf+1]     device: first row of facet f */
const uint16_t* ehdg_csr_dofmask(const ehdg_csr* csr);   /* u16 [nf]       device: active dof slots of facet f */
/* deterministic numeric fill: every entry gathers its <= 2 element contributions in ascending
 * element id (a sorted-key segmented reduction with the segments read off the topology). */
int ehdg_csr_numeric(ehdg_ctx* ctx, const ehdg_plan* plan, const ehdg_csr* csr, const ehdg_mesh* mesh,
                     const double* S, const double* g, double* vals, double* rhs, void* stream);

/* a5-a8 fused: ehdg_assemble_condense with the condensed blocks stored straight into the value array of the pattern
 * (convection_diffusion.py:357-383 -- what acd.Assemble() with condense=True does in one pass).  A block (f, g), f != g, has
 * exactly one contributing element and is written in place; the diagonal blocks and the right-hand side take their second
 * contribution through a side buffer, first + second as in ehdg_csr_numeric: vals and rhs are IDENTICAL to those of
 * ehdg_assemble_condense + ehdg_csr_numeric, which stay for callers that want the element blocks S_K (parity tests, the
 * Dirichlet-lifting variant).  W, z as in ehdg_assemble_condense; the pattern caches a per-element scatter table. */
int ehdg_assemble_condense_csr(ehdg_ctx* ctx, const ehdg_plan* plan, ehdg_csr* csr, const ehdg_mesh* mesh, const double* lam,
                               const uint8_t* elem_active, const uint8_t* facet_mask, double* vals, double* rhs, double* W,
                               double* z, int32_t* status, void* stream);

/* Inhomogeneous Dirichlet data (reference Python/debug_ehdg.py:210-214: gfu.components[1].Set(exact, BND); f -= A gfu;
 * gfu += A^-1 f).  ehdg_dirichlet_project: ubnd f64[nf][p+1] = L2 projection of `exact` onto the Legendre modes of every
 * boundary facet (bonus rule), 0 on interior facets.  The _bc variants of a8 / a10 move the prescribed dofs to the right-hand
 * side of the condensed system and put them back in the element back-solve; ubnd = null is the homogeneous case. */
int ehdg_dirichlet_project(ehdg_ctx* ctx, const ehdg_mesh* mesh, double* ubnd, void* stream);
int ehdg_csr_numeric_bc(ehdg_ctx* ctx, const ehdg_plan* plan, const ehdg_csr* csr, const ehdg_mesh* mesh, const double* S,
                        const double* g, const double* ubnd, double* vals, double* rhs, void* stream);
int ehdg_backsolve_bc(ehdg_ctx* ctx, const ehdg_plan* plan, const ehdg_csr* csr, const ehdg_mesh* mesh, const double* W,
                      const double* z, const double* xhat, const double* ubnd, double* u, void* stream);

/* a9  restarted GMRES with block-Jacobi (facet block) right preconditioning on the condensed system
 * (replaces acd.mat.Inverse(ba_active_dofs, "pardiso"), convection_diffusion.py:385-387). */
typedef struct {
    int32_t restart;        /* Krylov dimension m */
    int32_t max_iters;
    double rtol;            /* on ||D (b - A x)|| / ||D b||, D = diag(1 / max_j |a_rj|): the row-equilibrated residual (the rows of
                             * marked elements carry penalties 1e5..1e10 times those of the bulk; their rounding errors alone
                             * put a floor under the unscaled residual) */
    int32_t iters;          /* out */
    int32_t converged;      /* out */
    double rel_residual;    /* out: true residual at exit, in the norm of rtol */
    double total_seconds;   /* out: device time of the whole solve */
    int64_t spmv_calls;     /* out */
    int32_t dropped_dofs;   /* out: facet dofs found dependent inside their facet block (coefficient kept 0) */
    int32_t use_strip_block;/* in : 1 = facets of marked elements form one exactly solved preconditioner block */
    int32_t strip_dofs;     /* out */
    int32_t strip_blocks;   /* out */
    int32_t jacobi_sweeps;  /* in : s >= 1 block-Jacobi Richardson sweeps per preconditioner application (polynomial
                             *      block-Jacobi); s * restart should exceed the number of facet layers along a streamline */
    int32_t use_line_block; /* in : 1 = line-block Jacobi: the chains of ehdg_csr_auto_lines / ehdg_csr_set_lines are the
                             *      diagonal blocks of the preconditioner, each solved exactly (block-tridiagonal LU);
                             *      the matrix values are NOT overwritten in this mode */
    int32_t line_target_dofs; /* in : BFS levels of a chain are merged into blocks of at most this many dofs (0: one level per block) */
    int32_t line_chains;    /* out: chains of the line preconditioner (incl. the strip chain) */
    int32_t line_blocks;    /* out: blocks over all chains */
    int32_t line_max_block; /* out: dofs of the largest block */
    double setup_seconds;   /* out: device + host time of the preconditioner set-up (inside total_seconds) */
    double precond_bytes;   /* out: bytes of factor data one preconditioner application streams */
    double rel_residual_unscaled; /* out: ||b - A x|| / ||b|| of the returned x in the plain Euclidean norm (kept rows) */
    double drop_tol;        /* in : selection rule for (near-)dependent facet dofs: a dof whose pivot in the factorisation of its
                             * preconditioner block falls below drop_tol x the block's largest diagonal entry is removed from the
                             * system as test and trial function (coefficient 0, row ignored); 0 = default 1e-6.  The reference
                             * prunes with a distance test at 1e-5 that misses such dofs; dropped_dofs counts them */
    int32_t line_max_blocks;/* in : chains of the line preconditioner longer than this many blocks are cut into equal segments, each
                             *      solved exactly (0 = whole lines, < 0 = automatic: up to 4 segments per line while the GPU has fewer than
                             *      ~2400 chains).  Shorter chains = more, shorter sequential sweeps (the sweep is
                             *      latency-bound) at the price of a weaker preconditioner */
    int32_t reserved2;
} ehdg_gmres_info;
int ehdg_gmres(ehdg_ctx* ctx, ehdg_csr* csr, double* vals /* overwritten by A M^-1 unless use_line_block */,
               const double* rhs, double* x, ehdg_gmres_info* info, void* stream);
/* Same preconditioner (block-Jacobi, strip block, optional sweeps) with BiCGStab: short recurrence, no restart.
 * GMRES(m) stalls on the penalty-dominated facet systems of fine meshes; `restart` is ignored here. */
int ehdg_bicgstab(ehdg_ctx* ctx, ehdg_csr* csr, double* vals, const double* rhs, double* x, ehdg_gmres_info* info, void* stream);
/* DQGMRES(k), k = info->restart: GMRES with the orthogonalisation truncated to the last k Krylov vectors and a
 * progressively updated solution -- no restart, per-iteration cost of GMRES(k).  The method of choice for the
 * large facet systems, where information has to cross the mesh and restarted GMRES stagnates. */
int ehdg_dqgmres(ehdg_ctx* ctx, ehdg_csr* csr, double* vals, const double* rhs, double* x, ehdg_gmres_info* info, void* stream);
/* Caller-provided workspace of the Krylov solvers: ehdg_krylov_workspace_bytes gives the bytes the solver named
 * by `method` (0 gmres, 1 bicgstab, 2 dqgmres) needs for its vectors with the settings in `info`; ehdg_set_workspace hands
 * the context a device buffer (kept until replaced; ptr = null: back to per-call allocation).  Whatever does not fit is
 * allocated and freed inside the call.  The factors of the line / strip preconditioner are library-owned. */
int ehdg_krylov_workspace_bytes(const ehdg_ctx* ctx, const ehdg_csr* csr, const ehdg_gmres_info* info, int method, int64_t* bytes);
int ehdg_set_workspace(ehdg_ctx* ctx, void* device_ptr, int64_t bytes);
/* y = A x with the CSR SpMV kernel used inside GMRES (for tests and the bandwidth bench) */
int ehdg_spmv(ehdg_ctx* ctx, const ehdg_csr* csr, const double* vals, const double* x, double* y, void* stream);

/* ---- EDG path: enriched interior-penalty DG, Convection_Diffusion._solveEDG (convection_diffusion.py:42-219) ----
 * Spaces (:48-56): V = L2(order), Q_i = one P0 dof per element marked for enrichment i; u = u_0 + sum u_i psi_i.  Marking is
 * ehdg_mark (elem_mask), the plan is ehdg_plan_create with its facet_mask (general list = marked elements + neighbours).
 * ehdg_edg_prune   (:83-123)  mass Gram int u v of the enriched shapes with the bonus rule; enrichment k of a marked element is
 *                  dropped when the reference's `factor` <= theta (config['theta']).  elem_active u8[ne]; factors f64[ne][n_enrich]
 *                  (optional): -1 where the dof does not exist.  plan may be null (every element is visited).
 * ehdg_edg_alpha   (:129-163) lam[el] = lambda_max of (m, a + f f^T) with the bonus rule, incl. the carry-over of the previous
 *                  element's value where the eigenvalue is exactly zero (:158-161); the value the reference tabulates and
 *                  uses as alpha_stab.
 * ehdg_edg_csr_symbolic: pattern of acd.mat (:189-191, dgjumps=True): block row of element K = K and its facet neighbours in
 *                  ascending element number, rows / columns = active volume slots (polynomial modes, then surviving
 *                  enrichments).  The result is an ehdg_csr: ehdg_csr_auto_lines, ehdg_spmv and the Krylov solvers
 *                  (use_line_block = 1) take it as they take the condensed facet system.
 * ehdg_edg_assemble (:165-198) values of the block rows and the right-hand side: eps [grad u . grad v + pen [u][v] - {dn u}[v]
 *                  - {dn v}[u] (+ boundary terms)] - (b u) . grad v + (b.n) u_up [v],  pen = 4 alpha_stab p^2 / h^2.
 * ehdg_edg_unpack / ehdg_edg_l2_error (:205-210): u f64[ne][n_p + n_enrich] from the solution vector (inactive slots 0);
 *                  err2[2] = squared L2 and H1-seminorm errors against `exact` with the order 5 + bonus rule. */
/* Inhomogeneous Dirichlet data of the EDG method (reference Python/debug_edg.py:191-192): with enable != 0 ehdg_edg_assemble adds
 *   f += exact eps (alpha order^2 / h v - n . grad v) dS + b.n IfPos(b.n, 0, -exact) v dS
 * on the boundary facets (`exact` of the problem / ehdg_set_expression).  The bilinear form is the reference's: it has no
 * outflow term, the data enter through the Nitsche terms and the inflow flux only. */
int ehdg_edg_set_boundary_data(ehdg_ctx* ctx, int enable);
int ehdg_edg_prune(ehdg_ctx* ctx, const ehdg_plan* plan, const ehdg_mesh* mesh, const uint8_t* elem_mask, double theta,
                   uint8_t* elem_active, double* factors, void* stream);
int ehdg_edg_alpha(ehdg_ctx* ctx, const ehdg_plan* plan, const ehdg_mesh* mesh, const uint8_t* elem_active, double* lam,
                   int32_t* status, void* stream);
int ehdg_edg_alpha_carry(ehdg_ctx* ctx, int32_t ne, double* lam, void* stream);
int ehdg_edg_csr_symbolic(ehdg_ctx* ctx, const ehdg_mesh* mesh, const uint8_t* elem_active, ehdg_csr** out, void* stream);
int ehdg_edg_assemble(ehdg_ctx* ctx, const ehdg_mesh* mesh, const ehdg_csr* csr, const double* alpha_stab, const uint8_t* elem_active,
                      double* vals, double* rhs, int32_t* status, void* stream);
int ehdg_edg_unpack(ehdg_ctx* ctx, const ehdg_mesh* mesh, const ehdg_csr* csr, const uint8_t* elem_active, const double* x, double* u,
                    void* stream);
int ehdg_edg_l2_error(ehdg_ctx* ctx, const ehdg_mesh* mesh, const uint8_t* elem_active, const double* u, double* err2, void* stream);

/* Multi-GPU facet solve (one process per GPU).  Rows of the condensed system are split into contiguous
 * ranges; every rank holds the (replicated) matrix, iterates on its own rows, exchanges the SpMV halo rows
 * with its two neighbouring ranks (ncclSend/ncclRecv) and reduces every batch of dot products with one
 * ncclAllReduce.  libnccl.so.2 is resolved at run time.  id128: the 128-byte ncclUniqueId of rank 0
 * (ehdg_comm_unique_id), distributed by the caller (torch.distributed broadcast).
 * row_begin[nranks+1]; col_lo/col_hi[nranks]: column range referenced by each rank's rows. */
int ehdg_comm_unique_id(char* out128);
int ehdg_comm_init(ehdg_ctx* ctx, const char* id128, int nranks, int rank);
int ehdg_comm_set_partition(ehdg_ctx* ctx, const int64_t* row_begin, const int64_t* col_lo, const int64_t* col_hi);
int ehdg_comm_destroy(ehdg_ctx* ctx);

/* a10 back-solve U_K = z_K - W_K Uhat_K and L2 / H1 error against `exact`
 * (convection_diffusion.py:389-394).  u f64: polynomial elements [n_poly][n_p] then general [n_gen][nvmax].
 * err2 f64[2] (device): sum of squared L2 and H1-seminorm errors. */
int ehdg_backsolve(ehdg_ctx* ctx, const ehdg_plan* plan, const ehdg_csr* csr, const ehdg_mesh* mesh,
                   const double* W, const double* z, const double* xhat, double* u, void* stream);
int ehdg_l2_error(ehdg_ctx* ctx, const ehdg_plan* plan, const ehdg_mesh* mesh, const uint8_t* elem_active,
                  const double* u, double* err2, void* stream);
/* same, over the elements with elem_select[el] != 0 only (u8 [ne], device; null = all): the owned elements of a rank,
 * whose partial sums the caller adds over the ranks */
int ehdg_l2_error_select(ehdg_ctx* ctx, const ehdg_plan* plan, const ehdg_mesh* mesh, const uint8_t* elem_active,
                         const uint8_t* elem_select, const double* u, double* err2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EHDG_H */
