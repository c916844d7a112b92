// hy_cr3bp_match.hpp - host-side recognition of the CR3BP tape.
//
// hy_create receives the generic opcode tape (include/hy_cuda.h).  The tape of the reference's
// model.cr3bp (/root/reference/heyoka/expose_models.cpp:395-400), as lowered by the host
// (hy_b200/decompose.py), is the 24-op sequence below on the state (x, y, z, px, py, pz):
//
//    0,1   XA = x + cA, XB = x + cB                  LINCOMB  {x: 1}, {ONE: c}
//    2,3   y^2, z^2                                  SQUARE
//    4,5   x' = px + y, y' = py - x                  ADDSUB | SVD
//    6,7   XA^2, XB^2                                SQUARE
//    8,9   R1 = XA^2 + y^2 + z^2, R2 likewise        LINCOMB  (coefficients 1)
//   10,11  C1 = R1^(-3/2), C2 = R2^(-3/2)            POW
//   12,13  G1 = mA C1, G2 = mB C2                    LINCOMB  (one term)
//   14,15  G = gA C1 + gB C2, NG = nA C1 + nB C2     LINCOMB
//   16-19  XA G1, XB G2, y G, z NG                   MUL
//   20     px' = py - XA G1 - XB G2                  LINCOMB | SVD  (coefficients 1, -1, -1)
//   21     py' = -px - y G                           ADDSUB | SVD | NEGA | NEGB
//   22,23  z' = pz, pz' = z NG                       SVD
//
// The eight coefficients are free (they carry mu); everything else must match exactly, otherwise
// the tape stays on the interpreter.  No events, no runtime parameters, Taylor order up to the
// compiled ones (FP64: 20 - tol = eps - and 22 - tol = 1e-18; FP32: 9).
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/hy_cuda.h"
#include "hy_cr3bp_reg.cuh"

namespace hy {

struct CrbMatch {
    std::vector<double> imm; // CRB_NIMM entries: cA cB | mA mB | gA gB | nA nB
};

inline bool match_cr3bp(const hy_dims &d, const hy_op *ops, const hy_term *terms, int fp_bits, CrbMatch &out)
{
    // (FP64: orders above CrbPmax<double> take the order-22 build, see cr3bp_kernel_variant)
    const uint32_t pmax = fp_bits == 64 ? (uint32_t)CRB_PMAX_HI : (uint32_t)CrbPmax<float>::value;
    if (d.n_events || d.n_par || d.n_state != 6 || d.n_ops != 24 || d.order > pmax || d.order < 2) return false;
    const uint32_t P1 = d.order + 1;
    auto sv = [&](int i) { return HY_REF_JET | ((uint32_t)i * P1); };
    const uint32_t X = sv(0), Y = sv(1), Z = sv(2), PX = sv(3), PY = sv(4), PZ = sv(5);
    auto is = [&](int i, uint16_t opc, uint16_t flags) { return ops[i].opcode == opc && ops[i].flags == flags; };
    // terms of a LINCOMB: sources and (optionally) coefficients
    auto lin = [&](int i, uint16_t flags, std::initializer_list<uint32_t> src) -> const hy_term * {
        if (!is(i, HY_OP_LINCOMB, flags) || ops[i].n != src.size() || (uint64_t)ops[i].b + ops[i].n > d.n_terms)
            return nullptr;
        const hy_term *t = terms + ops[i].b;
        uint32_t q = 0;
        for (uint32_t s : src) {
            if (t[q].src != s || t[q].par >= 0) return nullptr;
            ++q;
        }
        return t;
    };
    auto un = [&](int i, uint16_t opc, uint32_t a) { return is(i, opc, 0) && ops[i].a == a; };
    out.imm.assign(CRB_NIMM, 0.0);
    // 0,1: XA, XB
    uint32_t XAB[2], SQ[2], RR[2], CC[2], GI[2], GG[2], T[4];
    for (int i = 0; i < 2; ++i) {
        const hy_term *t = lin(i, 0, {X, HY_REF_ONE});
        if (!t || t[0].coef != 1.0 || !(ops[i].dst & HY_REF_JET)) return false;
        out.imm[i] = t[1].coef;
        XAB[i] = ops[i].dst;
    }
    // 2,3: y^2, z^2; 6,7: XA^2, XB^2
    if (!un(2, HY_OP_SQUARE, Y) || !un(3, HY_OP_SQUARE, Z)) return false;
    if (!un(6, HY_OP_SQUARE, XAB[0]) || !un(7, HY_OP_SQUARE, XAB[1])) return false;
    SQ[0] = ops[6].dst;
    SQ[1] = ops[7].dst;
    // 4,5: x' = px + y, y' = py - x
    if (!is(4, HY_OP_ADDSUB, HY_OPF_SVD) || ops[4].dst != X || ops[4].a != PX || ops[4].b != Y) return false;
    if (!is(5, HY_OP_ADDSUB, HY_OPF_SVD | HY_OPF_NEGB) || ops[5].dst != Y || ops[5].a != PY || ops[5].b != X)
        return false;
    // 8,9: R = X^2 + y^2 + z^2; 10,11: C = R^(-3/2); 12,13: G_i = m_i C_i
    for (int i = 0; i < 2; ++i) {
        const hy_term *t = lin(8 + i, 0, {SQ[i], ops[2].dst, ops[3].dst});
        if (!t || t[0].coef != 1.0 || t[1].coef != 1.0 || t[2].coef != 1.0 || !(ops[8 + i].dst & HY_REF_JET))
            return false;
        RR[i] = ops[8 + i].dst;
        if (!un(10 + i, HY_OP_POW, RR[i]) || ops[10 + i].imm != -1.5 || !(ops[10 + i].dst & HY_REF_JET)) return false;
        CC[i] = ops[10 + i].dst;
        const hy_term *g = lin(12 + i, 0, {CC[i]});
        if (!g || !(ops[12 + i].dst & HY_REF_JET)) return false;
        out.imm[2 + i] = g[0].coef;
        GI[i] = ops[12 + i].dst;
    }
    // 14,15: G, NG
    for (int i = 0; i < 2; ++i) {
        const hy_term *t = lin(14 + i, 0, {CC[0], CC[1]});
        if (!t || !(ops[14 + i].dst & HY_REF_JET)) return false;
        out.imm[4 + 2 * i] = t[0].coef;
        out.imm[5 + 2 * i] = t[1].coef;
        GG[i] = ops[14 + i].dst;
    }
    // 16-19: products
    const uint32_t ma[4] = {XAB[0], XAB[1], Y, Z}, mb[4] = {GI[0], GI[1], GG[0], GG[1]};
    for (int i = 0; i < 4; ++i) {
        if (!is(16 + i, HY_OP_MUL, 0) || ops[16 + i].a != ma[i] || ops[16 + i].b != mb[i]) return false;
        T[i] = ops[16 + i].dst;
    }
    // 20: px' = py - T0 - T1
    {
        const hy_term *t = lin(20, HY_OPF_SVD, {PY, T[0], T[1]});
        if (!t || ops[20].dst != PX || t[0].coef != 1.0 || t[1].coef != -1.0 || t[2].coef != -1.0) return false;
    }
    // 21: py' = -px - T2
    if (!is(21, HY_OP_ADDSUB, HY_OPF_SVD | HY_OPF_NEGA | HY_OPF_NEGB) || ops[21].dst != PY || ops[21].a != PX ||
        ops[21].b != T[2])
        return false;
    // 22,23: z' = pz, pz' = T3
    if (!un(22, HY_OP_SVD, PZ) || ops[22].dst != Z || !un(23, HY_OP_SVD, T[3]) || ops[23].dst != PZ) return false;
    return true;
}

// kernel variant serving a matched tape at Taylor order p
inline uint32_t cr3bp_kernel_variant(uint32_t order, int fp_bits)
{
    return (fp_bits == 64 && order > (uint32_t)CrbPmax<double>::value) ? (uint32_t)CRB_VARIANT_P22 : (uint32_t)CRB_VARIANT;
}

} // namespace hy
