// ehdg_problem (C ABI) -> ehdg::Problem (kernel parameter block)
#pragma once
#include "../../include/ehdg.h"
#include "ehdg_general.h"

namespace ehdg {

inline bool problem_valid(const ehdg_problem& ep) {
    if (ep.order < 1 || ep.order > EHDG_MAX_ORDER) return false;
    if (ep.n_enrich < 0 || ep.n_enrich > EHDG_MAX_ENRICH) return false;
    if (ep.bonus_int < 0 || ep.bonus_int > 40) return false;
    if (!(ep.epsilon > 0.0)) return false;
    for (int k = 0; k < ep.n_enrich; ++k)
        if (ep.enrich_axis[k] != 0 && ep.enrich_axis[k] != 1) return false;
    return true;
}

inline Problem make_problem(const ehdg_problem& ep) {
    Problem pb;
    pb.p = ep.order;
    pb.n_p = (ep.order + 1) * (ep.order + 2) / 2;
    pb.nfl = ep.order + 1;
    pb.nenr = ep.n_enrich;
    pb.nb = pb.nfl + pb.nenr;
    pb.nvmax = pb.n_p + pb.nenr;
    pb.nfmax = 3 * pb.nb;
    pb.bonus = ep.bonus_int;
    pb.eps = ep.epsilon;
    pb.bx = ep.beta[0];
    pb.by = ep.beta[1];
    pb.vol_drop_tol = 1e-6;      // ehdg_set_drop_tol
    pb.edg_bnd_data = 0;         // ehdg_edg_set_boundary_data
    pb.xcoeff = pb.xexact = pb.xexact_dx = pb.xexact_dy = nullptr;      // ehdg_set_expression
    for (int k = 0; k < EHDG_MAX_ENRICH; ++k) {
        pb.enr_axis[k] = k < ep.n_enrich ? ep.enrich_axis[k] : 0;
        pb.enr_k[k] = k < ep.n_enrich ? ep.enrich_k[k] : 0.0;
    }
    for (int i = 0; i < 4; ++i) { pb.coeff.c[i] = ep.coeff.c[i]; pb.exact.c[i] = ep.exact.c[i]; }
    for (int i = 0; i < 2; ++i) { pb.coeff.k[i] = ep.coeff.k[i]; pb.exact.k[i] = ep.exact.k[i]; }
    return pb;
}

}  // namespace ehdg
