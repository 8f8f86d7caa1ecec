// Host-side quadrature rules: Gauss-Legendre on [0,1] and the
// Duffy-collapsed Gauss-Jacobi(1,0) x Gauss-Legendre rule on the reference triangle,
// floor(order/2)+1 points per direction.  Nodes by Newton iteration with deflation in long double.
#pragma once
#include <cmath>
#include <vector>

namespace ehdg {

// Jacobi P_n^(a,0) and derivative at z (long double)
inline void jacobi_a0(int n, long double a, long double z, long double* val, long double* der) {
    long double pm1 = 0.0L, p = 1.0L, dm1 = 0.0L, d = 0.0L;
    for (int j = 1; j <= n; ++j) {
        long double pn, dn;
        if (j == 1) {
            pn = 0.5L * ((a + 2.0L) * z + a);
            dn = 0.5L * (a + 2.0L);
        } else {
            const long double m = j - 1;
            const long double c0 = 2.0L * (m + 1.0L) * (m + a + 1.0L) * (2.0L * m + a);
            const long double c1 = (2.0L * m + a + 1.0L) * (2.0L * m + a + 2.0L) * (2.0L * m + a);
            const long double c2 = (2.0L * m + a + 1.0L) * a * a;
            const long double c3 = 2.0L * m * (m + a) * (2.0L * m + a + 2.0L);
            pn = ((c1 * z + c2) * p - c3 * pm1) / c0;
            dn = (c1 * p + (c1 * z + c2) * d - c3 * dm1) / c0;
        }
        pm1 = p; p = pn; dm1 = d; d = dn;
    }
    *val = p;
    *der = d;
}

// n-point Gauss-Jacobi(a,0) rule on [-1,1] for weight (1-z)^a, a in {0, 1}
inline void gauss_jacobi_a0(int n, int a, std::vector<long double>& x, std::vector<long double>& w) {
    x.assign(n, 0.0L);
    w.assign(n, 0.0L);
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 0; i < n; ++i) {
        long double z = -std::cos((2.0L * i + 1.0L) * pi / (2.0L * n + a));   // ascending guesses
        for (int it = 0; it < 100; ++it) {
            long double p, d;
            jacobi_a0(n, (long double)a, z, &p, &d);
            long double s = 0.0L;
            for (int j = 0; j < i; ++j) s += 1.0L / (z - x[j]);
            const long double dz = p / (d - s * p);
            z -= dz;
            if (std::fabs((double)dz) < 1e-19) break;
        }
        x[i] = z;
    }
    for (int i = 0; i < n; ++i) {
        long double p, d;
        jacobi_a0(n, (long double)a, x[i], &p, &d);
        // Gamma factors reduce to 1 for (a,0), a in {0,1}:  w = 2^(a+1) / ((1 - z^2) P_n'(z)^2)
        w[i] = (a == 0 ? 2.0L : 4.0L) / ((1.0L - x[i] * x[i]) * d * d);
    }
}

struct SegRule {
    std::vector<double> t, w;   // nodes in [0,1], weights summing to 1
};
struct TrigRule {
    std::vector<double> x, y, w;   // reference coordinates (lam0, lam1), weights summing to 1/2
};

inline int npts_1d(int order) { return order / 2 + 1; }

inline SegRule seg_rule(int order) {
    const int n = npts_1d(order);
    std::vector<long double> x, w;
    gauss_jacobi_a0(n, 0, x, w);
    SegRule r;
    for (int i = 0; i < n; ++i) {
        r.t.push_back((double)(0.5L * (x[i] + 1.0L)));
        r.w.push_back((double)(0.5L * w[i]));
    }
    return r;
}

inline TrigRule trig_rule(int order) {
    const int n = npts_1d(order);
    std::vector<long double> xj, wj, xl, wl;
    gauss_jacobi_a0(n, 1, xj, wj);
    gauss_jacobi_a0(n, 0, xl, wl);
    TrigRule r;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const long double xi = 0.5L * (xj[i] + 1.0L);
            const long double eta = 0.5L * (xl[j] + 1.0L);
            r.x.push_back((double)xi);
            r.y.push_back((double)(eta * (1.0L - xi)));
            r.w.push_back((double)(0.25L * wj[i] * 0.5L * wl[j]));
        }
    return r;
}

}  // namespace ehdg
