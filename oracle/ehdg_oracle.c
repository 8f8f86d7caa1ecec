/* CPU ORACLE, C restatement (test infrastructure, NOT product code; PARITY UNPINNED at the
 * matrix-entry level for the reasons given in oracle/ehdg_oracle.py).
 *
 * Restates, element by element with plain loops + OpenMP over elements, the quadrature
 * algorithm the reference runs through NGSolve:
 *   alpha_K      Python/convection_diffusion.py:314-351
 *   forms        Python/convection_diffusion.py:357-375      rhs :380-381
 *   enriched shapes  Python/enrichment_proxy.py:95-96,133-138 (volume), :161-162 (facet)
 * plus the static condensation S_K = A_FF - A_FV A_VV^-1 A_VF (NGSolve condense=True).
 * Basis and rules come in as TABLES built by the numpy oracle (symbolic Dubiner coefficients,
 * scipy Gauss rules), so this file shares no code with the CUDA library it checks.
 * Used for: parity tests at sizes the numpy oracle cannot reach, and the CPU baseline of bench.py.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -o oracle/_build/libehdg_oracle.so oracle/ehdg_oracle.c -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXP 6
#define MAXNP 28
#define MAXNV 30
#define MAXNB 9
#define MAXNF 27
#define MAXQ 1024

typedef struct {
    int p, n_p, nenr, bonus;
    double eps, bx, by;
    int enr_axis[2];
    double enr_k[2];
    double coeff_c[4], coeff_k[2];
    /* tables */
    const double* mono;      /* [n_p][p+1][p+1] monomial coefficients of the Dubiner functions in (X, Y) */
    int nq, nqf, nq0, nqf0;
    const double *vx, *vy, *vw, *st, *sw;          /* rules of order 2p + bonus */
    const double *vx0, *vy0, *vw0, *st0, *sw0;     /* rules of order 2p */
} orc_params;

static void layer(double k, double d, double* val, double* der) {
    const double e0 = exp(-k), E = exp(-k * d);
    *val = (1.0 - d) + (E - e0) / (e0 - 1.0);
    *der = 1.0 + k * E / (e0 - 1.0);
}

static double layer_expr(const double* c, const double* k, double dx, double dy) {
    double lx = 0, ly = 0, t;
    if (c[1] != 0 || c[3] != 0) layer(k[0], dx, &lx, &t);
    if (c[2] != 0 || c[3] != 0) layer(k[1], dy, &ly, &t);
    return c[0] + c[1] * lx + c[2] * ly + c[3] * lx * ly;
}

/* Dubiner values and (X,Y) derivatives from the monomial table */
static void dubiner(const orc_params* P, double X, double Y, double* v, double* dX, double* dY) {
    const int p = P->p, n1 = p + 1;
    double xp[MAXP + 1], yp[MAXP + 1];
    xp[0] = yp[0] = 1.0;
    for (int a = 1; a <= p; ++a) { xp[a] = xp[a - 1] * X; yp[a] = yp[a - 1] * Y; }
    for (int k = 0; k < P->n_p; ++k) {
        const double* C = P->mono + (size_t)k * n1 * n1;
        double s = 0, sx = 0, sy = 0;
        for (int a = 0; a <= p; ++a)
            for (int b = 0; b <= p; ++b) {
                const double c = C[a * n1 + b];
                if (c == 0.0) continue;
                s += c * xp[a] * yp[b];
                if (a > 0) sx += c * a * xp[a - 1] * yp[b];
                if (b > 0) sy += c * b * xp[a] * yp[b - 1];
            }
        v[k] = s; dX[k] = sx; dY[k] = sy;
    }
}

typedef struct {
    double P[3][2], det, adet, gl[3][2], omx[3], omy[3];
    int f0, f1, vn[3];
} geom_t;

static void geometry(const double* verts, const int* tri, geom_t* g) {
    for (int v = 0; v < 3; ++v) {
        g->vn[v] = tri[v];
        g->P[v][0] = verts[2 * tri[v]];
        g->P[v][1] = verts[2 * tri[v] + 1];
        g->omx[v] = 1.0 - g->P[v][0];
        g->omy[v] = 1.0 - g->P[v][1];
    }
    const double ax = g->P[0][0] - g->P[2][0], ay = g->P[0][1] - g->P[2][1];
    const double bx = g->P[1][0] - g->P[2][0], by = g->P[1][1] - g->P[2][1];
    g->det = ax * by - bx * ay;
    g->adet = fabs(g->det);
    g->gl[0][0] = by / g->det;  g->gl[0][1] = -bx / g->det;
    g->gl[1][0] = -ay / g->det; g->gl[1][1] = ax / g->det;
    g->gl[2][0] = -g->gl[0][0] - g->gl[1][0];
    g->gl[2][1] = -g->gl[0][1] - g->gl[1][1];
    int o[3] = {0, 1, 2};
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (g->vn[o[j]] < g->vn[o[i]]) { int t = o[i]; o[i] = o[j]; o[j] = t; }
    g->f0 = o[0];
    g->f1 = o[1];
}

/* shapes at local barycentrics l[3]: nv = n_p + nenr values */
static void vol_shapes(const orc_params* P, const geom_t* g, int volmask, const double* l, double* phi, double* gx, double* gy) {
    double v[MAXNP], dX[MAXNP], dY[MAXNP];
    dubiner(P, l[g->f0], l[g->f1], v, dX, dY);
    for (int i = 0; i < P->n_p; ++i) {
        phi[i] = v[i];
        gx[i] = dX[i] * g->gl[g->f0][0] + dY[i] * g->gl[g->f1][0];
        gy[i] = dX[i] * g->gl[g->f0][1] + dY[i] * g->gl[g->f1][1];
    }
    for (int k = 0; k < P->nenr; ++k) {
        double val = 0, der = 0;
        if ((volmask >> k) & 1) {
            const double* om = P->enr_axis[k] == 0 ? g->omx : g->omy;
            layer(P->enr_k[k], l[0] * om[0] + l[1] * om[1] + l[2] * om[2], &val, &der);
        }
        phi[P->n_p + k] = val;
        gx[P->n_p + k] = P->enr_axis[k] == 0 ? der : 0.0;
        gy[P->n_p + k] = P->enr_axis[k] == 0 ? 0.0 : der;
    }
}

static const int EDGE_A[3] = {2, 1, 0}, EDGE_B[3] = {0, 2, 1};   /* NGSolve ET_TRIG edges */

typedef struct { int a, b; double len, nx, ny, h; double xs; } edge_t;

static void edge_geometry(const geom_t* g, int e, edge_t* E) {
    E->a = EDGE_A[e]; E->b = EDGE_B[e];
    const double tx = g->P[E->b][0] - g->P[E->a][0], ty = g->P[E->b][1] - g->P[E->a][1];
    E->len = sqrt(tx * tx + ty * ty);
    double nx = ty / E->len, ny = -tx / E->len;
    const int c = 3 - E->a - E->b;
    if (nx * (g->P[E->a][0] - g->P[c][0]) + ny * (g->P[E->a][1] - g->P[c][1]) < 0) { nx = -nx; ny = -ny; }
    E->nx = nx; E->ny = ny;
    E->h = g->adet / E->len;
    E->xs = g->vn[E->a] < g->vn[E->b] ? 1.0 : -1.0;
}

static void facet_shapes(const orc_params* P, const geom_t* g, const edge_t* E, int facmask, double t, double* fs) {
    const double xi = E->xs * (2.0 * t - 1.0);
    fs[0] = 1.0;
    if (P->p >= 1) fs[1] = xi;
    for (int n = 1; n < P->p; ++n) fs[n + 1] = ((2.0 * n + 1.0) * xi * fs[n] - n * fs[n - 1]) / (n + 1.0);
    for (int j = 0; j < P->nenr; ++j) {
        double val = 0, der;
        if ((facmask >> j) & 1) {
            const double* om = P->enr_axis[j] == 0 ? g->omx : g->omy;
            layer(P->enr_k[j], (1.0 - t) * om[E->a] + t * om[E->b], &val, &der);
        }
        fs[P->p + 1 + j] = val;
    }
}

/* dense helpers ------------------------------------------------------------------------------ */
static int lu_solve(int n, double* A, int nrhs, double* B) {   /* A n x n, B n x nrhs, row-major; partial pivoting */
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double best = fabs(A[k * n + k]);
        for (int i = k + 1; i < n; ++i)
            if (fabs(A[i * n + k]) > best) { best = fabs(A[i * n + k]); piv = i; }
        if (best == 0.0) return -1;
        if (piv != k) {
            for (int j = 0; j < n; ++j) { double t = A[k * n + j]; A[k * n + j] = A[piv * n + j]; A[piv * n + j] = t; }
            for (int j = 0; j < nrhs; ++j) { double t = B[k * nrhs + j]; B[k * nrhs + j] = B[piv * nrhs + j]; B[piv * nrhs + j] = t; }
        }
        for (int i = k + 1; i < n; ++i) {
            const double f = A[i * n + k] / A[k * n + k];
            if (f == 0.0) continue;
            for (int j = k + 1; j < n; ++j) A[i * n + j] -= f * A[k * n + j];
            for (int j = 0; j < nrhs; ++j) B[i * nrhs + j] -= f * B[k * nrhs + j];
        }
    }
    for (int i = n - 1; i >= 0; --i)
        for (int j = 0; j < nrhs; ++j) {
            double s = B[i * nrhs + j];
            for (int k = i + 1; k < n; ++k) s -= A[i * n + k] * B[k * nrhs + j];
            B[i * nrhs + j] = s / A[i * n + i];
        }
    return 0;
}

/* largest eigenvalue of the symmetric pencil (M, A), A SPD: Cholesky + cyclic Jacobi */
static double pencil_max(int n, double* A, double* M) {
    double L[MAXNV * MAXNV], C[MAXNV * MAXNV];
    memset(L, 0, sizeof(double) * n * n);
    for (int k = 0; k < n; ++k) {
        double d = A[k * n + k];
        for (int j = 0; j < k; ++j) d -= L[k * n + j] * L[k * n + j];
        const int drop = !(d > 1e-13 * fabs(A[k * n + k]));
        L[k * n + k] = drop ? 0.0 : sqrt(d);
        for (int i = k + 1; i < n; ++i) {
            double s = A[i * n + k];
            for (int j = 0; j < k; ++j) s -= L[i * n + j] * L[k * n + j];
            L[i * n + k] = drop ? 0.0 : s / L[k * n + k];
        }
    }
    /* C = L^-1 M L^-T */
    for (int c = 0; c < n; ++c)
        for (int i = 0; i < n; ++i) {
            double s = M[i * n + c];
            for (int j = 0; j < i; ++j) s -= L[i * n + j] * C[j * n + c];
            C[i * n + c] = L[i * n + i] > 0 ? s / L[i * n + i] : 0.0;
        }
    for (int r = 0; r < n; ++r)
        for (int i = 0; i < n; ++i) {
            double s = C[r * n + i];
            for (int j = 0; j < i; ++j) s -= L[i * n + j] * M[r * n + j];
            M[r * n + i] = L[i * n + i] > 0 ? s / L[i * n + i] : 0.0;
        }
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) { const double s = 0.5 * (M[i * n + j] + M[j * n + i]); M[i * n + j] = M[j * n + i] = s; }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0, dia = 0;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) { if (i != j) off += M[i * n + j] * M[i * n + j]; else dia += M[i * n + i] * M[i * n + i]; }
        if (off <= 1e-30 * dia || off == 0.0) break;
        for (int pI = 0; pI < n - 1; ++pI)
            for (int q = pI + 1; q < n; ++q) {
                const double apq = M[pI * n + q];
                if (apq == 0.0) continue;
                const double th = (M[q * n + q] - M[pI * n + pI]) / (2.0 * apq);
                const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = M[k * n + pI], akq = M[k * n + q];
                    M[k * n + pI] = c * akp - s * akq;
                    M[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = M[pI * n + k], aqk = M[q * n + k];
                    M[pI * n + k] = c * apk - s * aqk;
                    M[q * n + k] = s * apk + c * aqk;
                }
            }
    }
    double best = M[0];
    for (int i = 1; i < n; ++i) if (M[i * n + i] > best) best = M[i * n + i];
    return best;
}

/* alpha eigenvalue of one element (conv_diff.py:314-351) */
static double alpha_element(const orc_params* P, const double* verts, const int* tri, int volmask, int exclude_const) {
    geom_t g;
    geometry(verts, tri, &g);
    const int nv = P->n_p + P->nenr;
    double a[MAXNV * MAXNV], m[MAXNV * MAXNV], f[MAXNV], phi[MAXNV], gx[MAXNV], gy[MAXNV];
    memset(a, 0, sizeof(a)); memset(m, 0, sizeof(m)); memset(f, 0, sizeof(f));
    for (int q = 0; q < P->nq0; ++q) {
        const double l[3] = {P->vx0[q], P->vy0[q], 1.0 - P->vx0[q] - P->vy0[q]};
        vol_shapes(P, &g, volmask, l, phi, gx, gy);
        const double w = P->vw0[q] * g.adet;
        for (int i = 0; i < nv; ++i) {
            f[i] += w * phi[i] / g.adet;
            for (int j = 0; j < nv; ++j) a[i * nv + j] += w * (gx[i] * gx[j] + gy[i] * gy[j]);
        }
    }
    for (int i = 0; i < nv; ++i) for (int j = 0; j < nv; ++j) a[i * nv + j] += f[i] * f[j];
    for (int e = 0; e < 3; ++e) {
        edge_t E;
        edge_geometry(&g, e, &E);
        for (int q = 0; q < P->nqf0; ++q) {
            double l[3] = {0, 0, 0};
            l[E.a] = 1.0 - P->st0[q]; l[E.b] = P->st0[q];
            vol_shapes(P, &g, volmask, l, phi, gx, gy);
            const double w = P->sw0[q] * E.len * E.h;
            double dn[MAXNV];
            for (int i = 0; i < nv; ++i) dn[i] = gx[i] * E.nx + gy[i] * E.ny;
            for (int i = 0; i < nv; ++i) for (int j = 0; j < nv; ++j) m[i * nv + j] += w * dn[i] * dn[j];
        }
    }
    int keep[MAXNV], n = 0;
    for (int i = 0; i < nv; ++i) {
        int ok = 1;
        if (i == 0 && exclude_const) ok = 0;
        if (i >= P->n_p && !((volmask >> (i - P->n_p)) & 1)) ok = 0;
        if (ok) keep[n++] = i;
    }
    double A2[MAXNV * MAXNV], M2[MAXNV * MAXNV];
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { A2[i * n + j] = a[keep[i] * nv + keep[j]]; M2[i * n + j] = m[keep[i] * nv + keep[j]]; }
    return pencil_max(n, A2, M2);
}

/* element blocks + condensation; padded layout nv = n_p + nenr, nf = 3 (p+1+nenr) */
static int condense_element(const orc_params* P, const double* verts, const int* tri, double lam, int volmask, const int* facmask,
                            double* S, double* gv, double* Afull) {
    geom_t g;
    geometry(verts, tri, &g);
    const int nv = P->n_p + P->nenr, nb = P->p + 1 + P->nenr, nf = 3 * nb, N = nv + nf;
    double* A = (double*)calloc((size_t)N * N + N, sizeof(double));
    double* F = A + (size_t)N * N;
    double phi[MAXNV], gx[MAXNV], gy[MAXNV], fs[MAXNB];
    for (int q = 0; q < P->nq; ++q) {
        const double l[3] = {P->vx[q], P->vy[q], 1.0 - P->vx[q] - P->vy[q]};
        vol_shapes(P, &g, volmask, l, phi, gx, gy);
        const double w = P->vw[q] * g.adet;
        const double dx = l[0] * g.omx[0] + l[1] * g.omx[1] + l[2] * g.omx[2];
        const double dy = l[0] * g.omy[0] + l[1] * g.omy[1] + l[2] * g.omy[2];
        const double cf = layer_expr(P->coeff_c, P->coeff_k, dx, dy);
        for (int i = 0; i < nv; ++i) {
            F[i] += w * cf * phi[i];
            const double bg = P->bx * gx[i] + P->by * gy[i];
            for (int j = 0; j < nv; ++j) A[i * N + j] += w * (P->eps * (gx[i] * gx[j] + gy[i] * gy[j]) - bg * phi[j]);
        }
    }
    const double alpha = 20.0 * lam;
    for (int e = 0; e < 3; ++e) {
        edge_t E;
        edge_geometry(&g, e, &E);
        const double bn = P->bx * E.nx + P->by * E.ny;
        const double pos = bn > 0 ? 1.0 : 0.0, neg = 1.0 - pos;
        const double tau = P->eps * alpha * (double)(P->p * P->p) / E.h;
        for (int q = 0; q < P->nqf; ++q) {
            const double t = P->st[q];
            double l[3] = {0, 0, 0};
            l[E.a] = 1.0 - t; l[E.b] = t;
            vol_shapes(P, &g, volmask, l, phi, gx, gy);
            facet_shapes(P, &g, &E, facmask[e], t, fs);
            const double w = P->sw[q] * E.len;
            double dn[MAXNV];
            for (int i = 0; i < nv; ++i) dn[i] = gx[i] * E.nx + gy[i] * E.ny;
            for (int i = 0; i < nv; ++i) {
                for (int j = 0; j < nv; ++j)
                    A[i * N + j] += w * ((tau + pos * bn) * phi[i] * phi[j] - P->eps * (phi[i] * dn[j] + dn[i] * phi[j]));
                for (int k = 0; k < nb; ++k) {
                    const int c = nv + e * nb + k;
                    A[i * N + c] += w * ((-tau + neg * bn) * phi[i] * fs[k] + P->eps * dn[i] * fs[k]);
                    A[c * N + i] += w * ((-tau - pos * bn - pos) * phi[i] * fs[k] + P->eps * dn[i] * fs[k]);
                }
            }
            for (int k = 0; k < nb; ++k)
                for (int k2 = 0; k2 < nb; ++k2)
                    A[(nv + e * nb + k) * N + nv + e * nb + k2] += w * (tau - neg * bn + pos) * fs[k] * fs[k2];
        }
    }
    if (Afull) memcpy(Afull, A, sizeof(double) * ((size_t)N * N + N));
    /* active sets */
    int iv[MAXNV], jf[MAXNF], niv = 0, njf = 0;
    for (int i = 0; i < nv; ++i) if (i < P->n_p || ((volmask >> (i - P->n_p)) & 1)) iv[niv++] = i;
    for (int e = 0; e < 3; ++e) for (int k = 0; k < nb; ++k) if (k <= P->p || ((facmask[e] >> (k - P->p - 1)) & 1)) jf[njf++] = e * nb + k;
    double Avv[MAXNV * MAXNV], B[MAXNV * (MAXNF + 1)];
    for (int i = 0; i < niv; ++i) {
        for (int j = 0; j < niv; ++j) Avv[i * niv + j] = A[iv[i] * N + iv[j]];
        for (int c = 0; c < njf; ++c) B[i * (njf + 1) + c] = A[iv[i] * N + nv + jf[c]];
        B[i * (njf + 1) + njf] = F[iv[i]];
    }
    const int rc = lu_solve(niv, Avv, njf + 1, B);
    memset(S, 0, sizeof(double) * nf * nf);
    memset(gv, 0, sizeof(double) * nf);
    for (int r = 0; r < njf; ++r) {
        for (int c = 0; c <= njf; ++c) {
            double s = (c < njf) ? A[(nv + jf[r]) * N + nv + jf[c]] : 0.0;
            for (int j = 0; j < niv; ++j) s -= A[(nv + jf[r]) * N + iv[j]] * B[j * (njf + 1) + c];
            if (c < njf) S[jf[r] * nf + jf[c]] = s; else gv[jf[r]] = s;
        }
    }
    free(A);
    return rc;
}

/* batch entry: elements el_list[0..n), per-element masks; lam is computed when compute_alpha != 0 */
int orc_batch(const orc_params* P, const double* verts, const int* tris, int n, const int* el_list, const unsigned char* volmask,
              const unsigned char* facmask /* [n][3] */, int compute_alpha, double* lam /* [n] */, double* S /* [n][nf*nf] or NULL */,
              double* g /* [n][nf] or NULL */, double* Afull /* [n][N*N+N] or NULL */, int do_condense, int nthreads) {
    const int nb = P->p + 1 + P->nenr, nf = 3 * nb, nv = P->n_p + P->nenr, N = nv + nf;
    int bad = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : bad)
    for (int i = 0; i < n; ++i) {
        const int el = el_list[i];
        const int fm[3] = {facmask[3 * i], facmask[3 * i + 1], facmask[3 * i + 2]};
        if (compute_alpha) lam[i] = alpha_element(P, verts, tris + 3 * el, volmask[i], el == 0);
        if (do_condense) {
            double* Sl = S ? S + (size_t)i * nf * nf : (double*)malloc(sizeof(double) * nf * nf);
            double* gl = g ? g + (size_t)i * nf : (double*)malloc(sizeof(double) * nf);
            if (condense_element(P, verts, tris + 3 * el, lam[i], volmask[i], fm, Sl, gl, Afull ? Afull + (size_t)i * ((size_t)N * N + N) : NULL)) bad++;
            if (!S) free(Sl);
            if (!g) free(gl);
        }
    }
    return bad;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
