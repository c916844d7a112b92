// hy_nbody_match.hpp - host-side recognition of N-body tapes.
//
// hy_create receives the generic opcode tape (include/hy_cuda.h).  For the tape
// of a Newtonian N-body system in Cartesian coordinates - what the reference's
// model.nbody() builds (/root/reference/heyoka/expose_models.cpp:237-272):
//
//   per pair (A,B), component c:   d_c = x_A,c - x_B,c            (ADDSUB)
//   per pair:                      r2 = sum_c d_c^2               (SUMSQ, n = 3)
//                                  w  = r2^(-3/2)                 (POW)
//                                  t_c = d_c * w                  (MULSH, n = 3)
//   per body b, component c:       v_b,c' = sum_q coef * t_c(pair q)   (LINCOMB | SVD, NB-1 terms)
//                                  x_b,c' = v_b,c                 (SVD)
//
// with state variables ordered (x, y, z, vx, vy, vz) per body - it fills the
// descriptor consumed by the register-resident kernel (hy_nbody_reg.cuh).
// Anything else (events, parameters, missing pairs, other exponents, more than
// NBR_MAXB bodies, orders above NBR_PMAX - NBR_LMAX for 6 bodies in FP64) is left to the tape interpreter.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "../../include/hy_cuda.h"
#include "hy_nbody_reg.cuh"

namespace hy {

// Host view of the immediate-table / column layout of hy_nbody_reg.cuh for the two group sizes (the device
// constants NBR_MAXB, NBR_NLANES, NBR_LANE0, NBR_OFF0, NBR_NIMM, NBR_TB0 with and without HY_NBR_G32).
struct NbrHostLayout {
    int maxb, lanes, lane0, off0, nimm, ws;
    explicit NbrHostLayout(bool g32)
        : maxb(g32 ? 8 : 6), lanes(g32 ? 32 : 16), lane0(maxb * NBR_CS), off0(lane0 + lanes), nimm(off0 + maxb * NBR_CS),
          ws(maxb * NBR_BS)
    {
    }
};

struct NbMatch {
    uint32_t nb = 0, n_pairs = 0;
    bool g32 = false; // 7 or 8 bodies: 32-lane groups, one trajectory per warp (built at hy_create time, HY_NBR_G32)
    bool has_par = false; // masses (acceleration coefficients) scaled by runtime parameters: the kernel is
                          // built at hy_create time with HY_NBR_PAR (hy_jit.hpp), not one of the precompiled ones
    std::vector<double> imm; // NBR_NIMM entries (layout in hy_nbody_reg.cuh)
};

// body counts with a compiled register-resident kernel (hy_nb3.cu ... hy_nb6.cu)
inline bool nbody_kernel_compiled(uint32_t nb) { return nb >= 3 && nb <= 6; }
// body counts served by a kernel built at hy_create time on 32-lane groups
inline bool nbody_kernel_g32(uint32_t nb) { return nb == 7 || nb == 8; }
// kernel variant for a matched tape of nb bodies at Taylor order p (0: none compiled)
inline uint32_t nbody_kernel_variant(uint32_t nb, uint32_t order, int fp_bits)
{
    if (nbody_kernel_g32(nb)) return order <= (uint32_t)NBR_PMAX ? nb : 0u;
    if (!nbody_kernel_compiled(nb)) return 0;
    if (order <= (uint32_t)NBR_PMAX) return nb;
    return (nb == 6 && fp_bits == 64 && order <= (uint32_t)NBR_LMAX) ? (uint32_t)NBR_VARIANT_P22 : 0u;
}

inline bool match_nbody(const hy_dims &d, const hy_op *ops, const hy_term *terms, NbMatch &out)
{
    const uint32_t P1 = d.order + 1, n = d.n_state;
    // (orders NBR_PMAX + 1 .. NBR_LMAX: only the 6-body FP64 build has a kernel - the caller checks)
    if (d.n_events || n % 6 || d.order > (uint32_t)NBR_LMAX || d.order < 2) return false;
    const uint32_t NB = n / 6;
    if (NB < 2 || NB > 8u) return false;
    const bool g32 = NB > 6;
    const NbrHostLayout HL(g32);
    const uint32_t NP = NB * (NB - 1) / 2;
    if (d.n_ops != 3 * NP + 3 * NP + 6 * NB) return false;
    // state variable index of a jet reference (or -1)
    auto state_var = [&](uint32_t ref) -> int {
        if (!(ref & HY_REF_JET) || ref == HY_REF_ONE) return -1;
        const uint32_t b = ref & 0x7fffffffu;
        return (b % P1 == 0 && b / P1 < n) ? (int)(b / P1) : -1;
    };
    struct Diff {
        int va, vb;
    };
    std::map<uint32_t, Diff> diffs;              // ADDSUB dst -> (var a, var b)
    std::map<uint32_t, int> sumsq_pair, pow_pair; // dst -> pair index
    std::map<uint32_t, std::pair<int, int>> tout; // MULSH output ref -> (pair, component)
    struct Pair {
        int a = -1, b = -1;
        uint32_t dref[3] = {0, 0, 0};
        bool has_pow = false, has_mul = false;
    };
    std::vector<Pair> pairs;
    std::vector<char> have_x(n, 0), have_v(n, 0);
    struct Term {
        int first;     // pair
        double second; // coefficient
        int par;       // runtime parameter multiplying it, or -1
    };
    std::vector<std::vector<Term>> acc(n); // v variable -> (pair, coef, par) in term order
    for (uint32_t i = 0; i < d.n_ops; ++i) {
        const hy_op &o = ops[i];
        if (o.flags & HY_OPF_EVENT) return false;
        switch (o.opcode) {
        case HY_OP_ADDSUB: {
            if ((o.flags & (HY_OPF_NEGA | HY_OPF_NEGB | HY_OPF_SVD)) != HY_OPF_NEGB) return false;
            const int va = state_var(o.a), vb = state_var(o.b);
            if (va < 0 || vb < 0 || !(o.dst & HY_REF_JET)) return false;
            diffs[o.dst] = Diff{va, vb};
        } break;
        case HY_OP_SUMSQ: {
            if (o.n != 3 || (o.flags & HY_OPF_SVD) || (uint64_t)o.b + 3 > d.n_terms) return false;
            Pair p;
            for (int c = 0; c < 3; ++c) {
                auto it = diffs.find(terms[o.b + c].src);
                if (it == diffs.end()) return false;
                const int va = it->second.va, vb = it->second.vb;
                if (va % 6 != c || vb % 6 != c) return false;
                if (c == 0) {
                    p.a = va / 6;
                    p.b = vb / 6;
                } else if (p.a != va / 6 || p.b != vb / 6)
                    return false;
                p.dref[c] = terms[o.b + c].src;
            }
            if (p.a == p.b) return false;
            sumsq_pair[o.dst] = (int)pairs.size();
            pairs.push_back(p);
        } break;
        case HY_OP_POW: {
            auto it = sumsq_pair.find(o.a);
            if (it == sumsq_pair.end() || o.imm != -1.5 || pairs[it->second].has_pow) return false;
            pairs[it->second].has_pow = true;
            pow_pair[o.dst] = it->second;
        } break;
        case HY_OP_MULSH: {
            auto it = pow_pair.find(o.a);
            if (it == pow_pair.end() || o.n != 3 || (uint64_t)o.b + 3 > d.n_terms) return false;
            Pair &p = pairs[it->second];
            if (p.has_mul) return false;
            p.has_mul = true;
            for (int c = 0; c < 3; ++c) {
                if (terms[o.b + c].src != p.dref[c]) return false;
                tout[terms[o.b + c].dst & 0x7fffffffu] = {it->second, c};
            }
        } break;
        case HY_OP_LINCOMB: {
            if (!(o.flags & HY_OPF_SVD) || o.n != NB - 1 || (uint64_t)o.b + o.n > d.n_terms) return false;
            const int v = state_var(o.dst);
            if (v < 0 || v % 6 < 3 || have_v[v]) return false;
            have_v[v] = 1;
            for (uint32_t q = 0; q < o.n; ++q) {
                const hy_term &t = terms[o.b + q];
                if (t.par >= (int32_t)d.n_par) return false;
                auto it = tout.find(t.src & 0x7fffffffu);
                if (it == tout.end() || it->second.second != v % 6 - 3) return false;
                const Pair &p = pairs[it->second.first];
                if (p.a != v / 6 && p.b != v / 6) return false;
                acc[v].push_back(Term{it->second.first, t.coef, t.par});
            }
        } break;
        case HY_OP_SVD: {
            const int x = state_var(o.dst), v = state_var(o.a);
            if (x < 0 || x % 6 >= 3 || v != x + 3 || have_x[x]) return false;
            have_x[x] = 1;
        } break;
        default: return false;
        }
    }
    if (pairs.size() != NP || NP > (uint32_t)HL.lanes) return false;
    for (const Pair &p : pairs)
        if (!p.has_pow || !p.has_mul) return false;
    for (uint32_t b = 0; b < NB; ++b)
        for (int c = 0; c < 3; ++c)
            if (!have_x[6 * b + c] || !have_v[6 * b + 3 + c]) return false;
    // every body: the three components must list the same pairs with the same coefficients
    out = NbMatch();
    out.nb = NB;
    out.n_pairs = NP;
    out.g32 = g32;
    out.imm.assign(HL.nimm, 0.0);
    out.has_par = d.n_par != 0; // (parameters that no term uses still travel with the trajectory)
    std::vector<int> qa(NP, -1), qb(NP, -1);
    for (uint32_t b = 0; b < NB; ++b) {
        const auto &r0 = acc[6 * b + 3];
        for (int c = 1; c < 3; ++c) {
            const auto &rc = acc[6 * b + 3 + c];
            for (uint32_t q = 0; q + 1 < NB; ++q)
                if (rc[q].first != r0[q].first || rc[q].second != r0[q].second || rc[q].par != r0[q].par) return false;
        }
        for (uint32_t q = 0; q + 1 < NB; ++q) {
            const int pr = r0[q].first;
            int &slot = pairs[pr].a == (int)b ? qa[pr] : qb[pr];
            if (slot >= 0) return false; // the same pair twice in one sum
            slot = (int)q;
            out.imm[b * NBR_CS + q] = r0[q].second;
            // pair slot (= lane) whose products feed this term; second word: parameter index + 1 (0: none)
            const uint32_t src[2] = {(uint32_t)pr, (uint32_t)(r0[q].par + 1)};
            if (r0[q].par >= 0) out.has_par = true;
            std::memcpy(&out.imm[HL.off0 + b * NBR_CS + q], src, 8);
        }
    }
    for (uint32_t s = 0; s < (uint32_t)HL.lanes; ++s) {
        const uint32_t pr = s < NP ? s : 0; // idle lanes mirror pair 0
        if (qa[pr] < 0 || qb[pr] < 0) return false;
        const uint32_t rec[2] = {(uint32_t)pairs[pr].a, (uint32_t)pairs[pr].b}; // lane record
        std::memcpy(&out.imm[HL.lane0 + s], rec, 8);
    }
    return true;
}

} // namespace hy
