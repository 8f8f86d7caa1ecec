// Shape functions of the (E)HDG hot path, usable from host and device code.
//
// Restates NGSolve conventions that the reference relies on but does not contain
// (SURVEY.md Appendix A.3/A.4):
//   * L2(order=p) on a triangle: Dubiner basis on vertex-sorted barycentrics
//       phi_ij = P_i^scaled(Y - Z; 1 - X) * P_j^(2i+1,0)(2X - 1),  Z = 1 - X - Y,  i-major
//   * FacetFESpace(order=p) on an edge: Legendre polynomials in xi = lam_hi - lam_lo
//   * enrichment functions of run.py:26-27   p(s) = s + (e^{k(s-1)} - e^{-k}) / (e^{-k} - 1)
// All recurrences carry first derivatives along (forward mode), no tables involved.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define EHDG_HD __host__ __device__ __forceinline__
#else
#define EHDG_HD inline
#endif

#define EHDG_MAX_ORDER 6
#define EHDG_MAX_NP 28          // (6+1)(6+2)/2
#define EHDG_MAX_ENRICH 2

namespace ehdg {

// value and partial derivatives w.r.t. the two sorted barycentrics (X, Y)
struct Dual2 {
    double v, dx, dy;
};
EHDG_HD Dual2 d2(double v, double dx, double dy) { Dual2 r; r.v = v; r.dx = dx; r.dy = dy; return r; }
EHDG_HD Dual2 operator+(Dual2 a, Dual2 b) { return d2(a.v + b.v, a.dx + b.dx, a.dy + b.dy); }
EHDG_HD Dual2 operator-(Dual2 a, Dual2 b) { return d2(a.v - b.v, a.dx - b.dx, a.dy - b.dy); }
EHDG_HD Dual2 operator*(Dual2 a, Dual2 b) {
    return d2(a.v * b.v, a.dx * b.v + a.v * b.dx, a.dy * b.v + a.v * b.dy);
}
EHDG_HD Dual2 operator*(double s, Dual2 a) { return d2(s * a.v, s * a.dx, s * a.dy); }

// n_p Dubiner functions and their (X, Y) derivatives at one point.
//   v[k], dX[k], dY[k],  k = 0 .. (p+1)(p+2)/2 - 1   (i-major: i = 0..p, j = 0..p-i)
EHDG_HD void dubiner_eval(int p, double X, double Y, double* v, double* dX, double* dY) {
    const Dual2 x = d2(2.0 * Y + X - 1.0, 1.0, 2.0);   // Y - Z
    const Dual2 t = d2(1.0 - X, -1.0, 0.0);            // Y + Z
    const Dual2 z = d2(2.0 * X - 1.0, 2.0, 0.0);
    const Dual2 tt = t * t;
    Dual2 Lm1 = d2(0.0, 0.0, 0.0), L = d2(1.0, 0.0, 0.0);   // scaled Legendre P_{i-1}, P_i
    int k = 0;
    for (int i = 0; i <= p; ++i) {
        if (i == 1) {
            Lm1 = L;
            L = x;
        } else if (i > 1) {
            // i P_i = (2i-1) x P_{i-1} - (i-1) t^2 P_{i-2}
            const double n = (double)(i - 1);
            Dual2 Ln = (1.0 / (n + 1.0)) * ((2.0 * n + 1.0) * (x * L) - n * (tt * Lm1));
            Lm1 = L;
            L = Ln;
        }
        // Jacobi P_j^(a,0)(z), a = 2i+1
        const double a = 2.0 * i + 1.0;
        Dual2 Jm1 = d2(0.0, 0.0, 0.0), J = d2(1.0, 0.0, 0.0);
        for (int j = 0; j <= p - i; ++j) {
            if (j == 1) {
                Jm1 = J;
                J = 0.5 * ((a + 2.0) * z + d2(a, 0.0, 0.0));
            } else if (j > 1) {
                const double n = (double)(j - 1);
                const double c0 = 2.0 * (n + 1.0) * (n + a + 1.0) * (2.0 * n + a);
                const double c1 = (2.0 * n + a + 1.0) * (2.0 * n + a + 2.0) * (2.0 * n + a);
                const double c2 = (2.0 * n + a + 1.0) * a * a;
                const double c3 = 2.0 * n * (n + a) * (2.0 * n + a + 2.0);
                Dual2 Jn = (1.0 / c0) * ((c1 * z + d2(c2, 0.0, 0.0)) * J - c3 * Jm1);
                Jm1 = J;
                J = Jn;
            }
            const Dual2 f = L * J;
            v[k] = f.v;
            dX[k] = f.dx;
            dY[k] = f.dy;
            ++k;
        }
    }
}

// Legendre P_0..P_p at xi in [-1, 1]
EHDG_HD void legendre_eval(int p, double xi, double* mu) {
    mu[0] = 1.0;
    if (p >= 1) mu[1] = xi;
    for (int n = 1; n < p; ++n)
        mu[n + 1] = ((2.0 * n + 1.0) * xi * mu[n] - n * mu[n - 1]) / (n + 1.0);
}

// Boundary-layer function of run.py:26-27 written in the distance d = 1 - s to the outflow
// side (d >= 0 on the unit square), which keeps the exponent k*d free of cancellation:
//   L(s)  = s + (e^{k(s-1)} - e^{-k}) / (e^{-k} - 1) = (1 - d) + (e^{-k d} - e^{-k}) / (e^{-k} - 1)
//   L'(s) = 1 + k e^{-k d} / (e^{-k} - 1)
EHDG_HD void layer_eval(double k, double d, double* val, double* der) {
    const double e0 = exp(-k);
    const double E = exp(-k * d);
    const double den = e0 - 1.0;
    *val = (1.0 - d) + (E - e0) / den;
    *der = 1.0 + k * E / den;
}

// Same function with the per-problem constants hoisted: e0 = exp(-k), inv = 1 / (e0 - 1).
// exp(-k d) underflows to exactly 0 beyond k d = 745.2, so the exponential is skipped there (bit-identical).
struct LayerConst {
    double k, e0, inv;
};
EHDG_HD LayerConst layer_const(double k) {
    LayerConst c;
    c.k = k;
    c.e0 = exp(-k);
    c.inv = 1.0 / (c.e0 - 1.0);
    return c;
}
EHDG_HD double layer_val(const LayerConst& c, double d) {
    const double kd = c.k * d;
    const double E = kd > 746.0 ? 0.0 : exp(-kd);
    return (1.0 - d) + (E - c.e0) * c.inv;
}

// f(x, y) = c0 + c1 L(k0; x) + c2 L(k1; y) + c3 L(k0; x) L(k1; y), given dx = 1 - x, dy = 1 - y.
// Covers coeff = beta1 p(x) + beta0 q(y) (run.py:30) and exact = p(x) q(y) (run.py:29).
struct LayerExpr {
    double c[4];
    double k[2];
};
EHDG_HD double layer_expr_eval(const LayerExpr& f, double dx, double dy) {
    double lx = 0.0, ly = 0.0, tmp;
    if (f.c[1] != 0.0 || f.c[3] != 0.0) layer_eval(f.k[0], dx, &lx, &tmp);
    if (f.c[2] != 0.0 || f.c[3] != 0.0) layer_eval(f.k[1], dy, &ly, &tmp);
    return f.c[0] + f.c[1] * lx + f.c[2] * ly + f.c[3] * lx * ly;
}
EHDG_HD void layer_expr_grad(const LayerExpr& f, double dx, double dy, double* val, double* gx, double* gy) {
    double lx = 0.0, ly = 0.0, px = 0.0, py = 0.0;
    layer_eval(f.k[0], dx, &lx, &px);
    layer_eval(f.k[1], dy, &ly, &py);
    *val = f.c[0] + f.c[1] * lx + f.c[2] * ly + f.c[3] * lx * ly;
    *gx = f.c[1] * px + f.c[3] * px * ly;
    *gy = f.c[2] * py + f.c[3] * lx * py;
}

}  // namespace ehdg
