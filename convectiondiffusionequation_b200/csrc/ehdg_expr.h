// Generic coefficient functions on the device: config['coeff'] / config['exact'] of the reference are arbitrary NGSolve
// CoefficientFunction trees of x, y (run.py:26-30, :44).  Here an expression (sympy on the host, coefficients.py) is compiled
// to a POSTFIX PROGRAM that every kernel evaluates with a small register stack: no run-time compilation, no function
// pointers, one code path for every expression.  The built-in boundary-layer family (LayerExpr, evaluated through the
// cancellation-free distances 1 - x, 1 - y) stays the default and the fast path of the right-hand side.
#pragma once
#include "ehdg_basis.h"

namespace ehdg {

#define EHDG_EXPR_MAX 128       /* instructions per program */
#define EHDG_EXPR_STACK 16

enum ExprOp {
    EX_END = 0, EX_CONST, EX_X, EX_Y, EX_ADD, EX_SUB, EX_MUL, EX_DIV, EX_NEG, EX_EXP, EX_LOG, EX_SIN, EX_COS, EX_TAN, EX_SQRT,
    EX_POWI /* integer exponent in arg */, EX_POW, EX_TANH, EX_ABS, EX_ATAN, EX_SIGN
};

struct ExprProg {
    int n;
    int op[EHDG_EXPR_MAX];
    double arg[EHDG_EXPR_MAX];
};

EHDG_HD double expr_eval(const ExprProg* __restrict__ p, double x, double y) {
    double st[EHDG_EXPR_STACK];
    int sp = 0;
    for (int i = 0; i < p->n; ++i) {
        switch (p->op[i]) {
            case EX_CONST: st[sp++] = p->arg[i]; break;
            case EX_X: st[sp++] = x; break;
            case EX_Y: st[sp++] = y; break;
            case EX_ADD: --sp; st[sp - 1] += st[sp]; break;
            case EX_SUB: --sp; st[sp - 1] -= st[sp]; break;
            case EX_MUL: --sp; st[sp - 1] *= st[sp]; break;
            case EX_DIV: --sp; st[sp - 1] /= st[sp]; break;
            case EX_NEG: st[sp - 1] = -st[sp - 1]; break;
            case EX_EXP: st[sp - 1] = exp(st[sp - 1]); break;
            case EX_LOG: st[sp - 1] = log(st[sp - 1]); break;
            case EX_SIN: st[sp - 1] = sin(st[sp - 1]); break;
            case EX_COS: st[sp - 1] = cos(st[sp - 1]); break;
            case EX_TAN: st[sp - 1] = tan(st[sp - 1]); break;
            case EX_SQRT: st[sp - 1] = sqrt(st[sp - 1]); break;
            case EX_TANH: st[sp - 1] = tanh(st[sp - 1]); break;
            case EX_ABS: st[sp - 1] = fabs(st[sp - 1]); break;
            case EX_ATAN: st[sp - 1] = atan(st[sp - 1]); break;
            case EX_SIGN: st[sp - 1] = st[sp - 1] > 0.0 ? 1.0 : (st[sp - 1] < 0.0 ? -1.0 : 0.0); break;
            case EX_POWI: {
                int e = (int)p->arg[i];
                const bool inv = e < 0;
                if (inv) e = -e;
                double b = st[sp - 1], r = 1.0;
                while (e) { if (e & 1) r *= b; b *= b; e >>= 1; }
                st[sp - 1] = inv ? 1.0 / r : r;
                break;
            }
            case EX_POW: --sp; st[sp - 1] = pow(st[sp - 1], st[sp]); break;
            default: break;
        }
    }
    return sp > 0 ? st[sp - 1] : 0.0;
}

}  // namespace ehdg
