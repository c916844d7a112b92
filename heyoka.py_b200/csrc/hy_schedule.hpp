// hy_schedule.hpp - host-side scheduling of the ABI tape into per-lane programs.
//
// The ABI tape (include/hy_cuda.h) lists the elementary ops of one order sweep
// in dependency order.  A group of G lanes cooperates on one trajectory, so the
// ops must be distributed over the lanes with as few group synchronisations as
// possible.  This file
//   1. finds the same-sweep dependencies between ops,
//   2. merges every op whose consumers all live in one cluster into that
//      cluster (a cluster is executed start-to-end by ONE lane, so its internal
//      dependencies need no synchronisation; for the N-body problem a cluster is
//      one body pair: 3 differences, sum of squares, pow(-3/2), 3 products),
//   3. levels the cluster DAG into PHASES (one group sync per phase),
//   4. packs clusters with identical opcode signatures into rows of G lanes so
//      that the lanes of a warp run the same opcode at the same time,
//   5. emits the program in a lane-interleaved (ELL) layout: op slot j of lane s
//      sits at ops[(slot)*G + s] and term c of lane s at terms[c*G + s], so every
//      tape access of a warp is bank-conflict free (or a broadcast).
//
// This replaces the instruction scheduling LLVM performs for the reference's
// JIT-compiled stepper ([UPSTREAM] heyoka taylor_adaptive_batch ctor, called
// from /root/reference/heyoka/expose_batch_integrators.cpp:166-208).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <map>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/hy_cuda.h"
#include "hy_devprog.h"

namespace hy {

struct Program {
    uint32_t G = 1;
    uint32_t n_phases = 0;
    std::vector<uint32_t> phase_slot; // [n_phases + 1] slot ranges
    std::vector<DOp> ops;             // [n_slots * G]
    std::vector<DTerm> terms;         // [n_tslots * G]
    std::vector<double> imm;          // immediates
    uint32_t n_slots = 0, n_tslots = 0;
    // device workspace layout (per trajectory column)
    uint32_t ws_len = 0, par_off = 0, one_off = 0;
    uint32_t n_spill = 0;              // state jets kept in the global scratch
    uint32_t evt_bytes = 0;            // event tape of the register-resident kernels (hy_evtape.cuh)
    std::vector<uint32_t> state_row;   // [n_state] device row: jet base (resident) or ping-pong base (spilled)
    std::vector<int32_t> state_spill;  // [n_state] spill slot or -1
    std::vector<uint32_t> ev_ref;      // [n_events] remapped event jet references
    // statistics
    uint32_t n_clusters = 0;
    double lane_utilisation = 0; // useful op slots / (slots * G), cost weighted
};

namespace detail {

inline bool is_term_op(uint16_t oc) { return oc == HY_OP_LINCOMB || oc == HY_OP_SUMSQ || oc == HY_OP_MULSH; }

// Rough relative cost of an op at a mid order (used only for balancing).
inline double op_cost(const hy_op &o)
{
    const double k = 10;
    switch (o.opcode) {
    case HY_OP_LINCOMB: return 4 + 3.0 * o.n;
    case HY_OP_ADDSUB: return 5;
    case HY_OP_MUL: return 8 + 3 * k;
    case HY_OP_SQUARE: return 8 + 1.5 * k;
    case HY_OP_SUMSQ: return 8 + 1.5 * k * o.n;
    case HY_OP_MULSH: return 8 + (1.0 + o.n) * k + o.n * k * 0.5;
    case HY_OP_DIV: return 10 + 3 * k;
    case HY_OP_POW:
    case HY_OP_SQRT: return 12 + 5 * k;
    case HY_OP_EXP:
    case HY_OP_LOG:
    case HY_OP_INTG: return 10 + 4 * k;
    case HY_OP_SINCOS: return 12 + 6 * k;
    case HY_OP_TIME: return 3;
    case HY_OP_SVD: return 4;
    default: return 1;
    }
}

struct UF {
    std::vector<int> p;
    explicit UF(int n) : p(n) { std::iota(p.begin(), p.end(), 0); }
    int find(int x)
    {
        while (p[x] != x) x = p[x] = p[p[x]];
        return x;
    }
};

} // namespace detail

// Build the per-lane program.  Returns an empty string on success, else an
// error message.
inline std::string build_program(const hy_dims &d, const hy_op *ops, const hy_term *terms, const uint32_t *ev_ref_in,
                                 uint32_t G, bool spill_ok, Program &out, bool pair_fusion = true)
{
    using namespace detail;
    const uint32_t n_ops = d.n_ops;
    const uint32_t P1 = d.order + 1;

    // ---- 1. producers of rows written in the same sweep ----
    std::map<uint32_t, int> prod; // row base -> op
    auto base_of = [](uint32_t ref) { return ref & 0x7fffffffu; };
    for (uint32_t i = 0; i < n_ops; ++i) {
        const hy_op &o = ops[i];
        if (o.opcode >= HY_OP_COUNT) return "invalid opcode in the tape";
        if (o.opcode == HY_OP_SVD || (o.flags & HY_OPF_SVD)) continue; // writes order k+1: next sweep
        if (o.opcode == HY_OP_MULSH) {
            for (uint32_t j = 0; j < o.n; ++j) prod[base_of(terms[o.b + j].dst)] = (int)i;
        } else {
            prod[base_of(o.dst)] = (int)i;
            if (o.opcode == HY_OP_SINCOS) prod[base_of(o.dst2)] = (int)i;
        }
    }
    std::vector<std::vector<int>> deps(n_ops), cons(n_ops);
    auto add_dep = [&](uint32_t i, uint32_t ref) {
        if (ref == HY_REF_ONE) return;
        auto it = prod.find(base_of(ref));
        if (it != prod.end() && it->second != (int)i) {
            if (std::find(deps[i].begin(), deps[i].end(), it->second) == deps[i].end()) {
                deps[i].push_back(it->second);
                cons[it->second].push_back((int)i);
            }
        }
    };
    for (uint32_t i = 0; i < n_ops; ++i) {
        const hy_op &o = ops[i];
        if (is_term_op(o.opcode)) {
            if ((uint64_t)o.b + o.n > d.n_terms) return "term range out of bounds";
            for (uint32_t j = 0; j < o.n; ++j) add_dep(i, terms[o.b + j].src);
            if (o.opcode == HY_OP_MULSH) add_dep(i, o.a);
        } else if (o.opcode != HY_OP_TIME) {
            add_dep(i, o.a);
            if (o.opcode == HY_OP_MUL || o.opcode == HY_OP_DIV || o.opcode == HY_OP_ADDSUB || o.opcode == HY_OP_INTG)
                add_dep(i, o.b);
        }
    }
    for (uint32_t i = 0; i < n_ops; ++i)
        for (int j : deps[i])
            if (j > (int)i) return "the tape is not in dependency order";

    // ---- 2. clusters ----
    UF uf((int)n_ops);
    for (bool changed = true; changed;) {
        changed = false;
        for (uint32_t j = 0; j < n_ops; ++j) {
            if (cons[j].empty()) continue;
            const int c0 = uf.find(cons[j][0]);
            bool same = true;
            for (int c : cons[j]) same = same && uf.find(c) == c0;
            if (same && uf.find((int)j) != c0) {
                uf.p[uf.find((int)j)] = c0;
                changed = true;
            }
        }
    }
    std::map<int, std::vector<int>> members; // root -> ops in tape order
    for (uint32_t i = 0; i < n_ops; ++i) members[uf.find((int)i)].push_back((int)i);
    std::vector<std::vector<int>> clusters;
    std::vector<int> cl_of(n_ops);
    for (auto &kv : members) {
        for (int i : kv.second) cl_of[i] = (int)clusters.size();
        clusters.push_back(kv.second);
    }
    const int NC = (int)clusters.size();

    // ---- 2b. superinstruction fusion: recognise pair-interaction clusters ----
    //   ADDSUB x n, SUMSQ(n) over their outputs, POW of the sum, MULSH(n) of the
    //   differences with the power.  Arithmetic (and summation order) is identical
    //   to the unfused ops; only dispatch and sync overhead is saved.
    std::vector<char> fused(clusters.size(), 0);
    const bool allow_fuse = pair_fusion && std::getenv("HY_CUDA_NO_PAIR_FUSION") == nullptr;
    for (size_t c = 0; allow_fuse && c < clusters.size(); ++c) {
        const auto &m = clusters[c];
        if (m.size() < 5) continue;
        const size_t nd = m.size() - 3;
        if (nd < 1 || nd > 3) continue;
        bool ok = true;
        for (size_t i = 0; i < nd && ok; ++i) {
            const hy_op &o = ops[m[i]];
            ok = o.opcode == HY_OP_ADDSUB && !(o.flags & (HY_OPF_SVD | HY_OPF_EVENT)) && (o.dst & HY_REF_JET);
        }
        if (!ok) continue;
        const hy_op &sq = ops[m[nd]], &pw = ops[m[nd + 1]], &ms = ops[m[nd + 2]];
        ok = sq.opcode == HY_OP_SUMSQ && sq.n == nd && pw.opcode == HY_OP_POW && ms.opcode == HY_OP_MULSH &&
             ms.n == nd && !((sq.flags | pw.flags | ms.flags) & (HY_OPF_SVD | HY_OPF_EVENT)) &&
             (sq.dst & HY_REF_JET) && pw.a == sq.dst && ms.a == pw.dst;
        for (size_t i = 0; i < nd && ok; ++i) {
            ok = terms[sq.b + i].src == ops[m[i]].dst && terms[ms.b + i].src == ops[m[i]].dst;
        }
        if (ok) fused[c] = 1;
    }

    // ---- 3. phases = levels of the cluster DAG ----
    std::vector<int> lvl(NC, 0);
    std::vector<char> has_pred(NC, 0), has_succ(NC, 0);
    // clusters are numbered by their root op id, not topologically: iterate to a fixed point
    for (bool changed = true; changed;) {
        changed = false;
        for (uint32_t i = 0; i < n_ops; ++i)
            for (int j : deps[i]) {
                const int ci = cl_of[i], cj = cl_of[j];
                if (ci == cj) continue;
                has_pred[ci] = 1;
                has_succ[cj] = 1;
                if (lvl[ci] < lvl[cj] + 1) {
                    lvl[ci] = lvl[cj] + 1;
                    changed = true;
                }
            }
    }
    int n_ph = 0;
    for (int c = 0; c < NC; ++c) n_ph = std::max(n_ph, lvl[c] + 1);
    // Isolated clusters (e.g. x' = v recurrences) go where they disturb least: the last phase.
    for (int c = 0; c < NC; ++c)
        if (!has_pred[c] && !has_succ[c]) lvl[c] = n_ph - 1;

    // ---- 4. rows of equal-signature clusters ----
    auto signature = [&](int c) {
        std::vector<uint32_t> s;
        if (fused[c]) {
            s.push_back(((uint32_t)DOP_PAIR << 16) | (uint32_t)(clusters[c].size() - 3));
            return s;
        }
        for (int i : clusters[c]) s.push_back(((uint32_t)ops[i].opcode << 16) | (ops[i].n & 0xffffu));
        return s;
    };
    // program length of a cluster in op slots
    auto cl_len = [&](int c) { return fused[c] ? (size_t)1 : clusters[c].size(); };
    auto cl_cost = [&](int c) {
        double t = 0;
        for (int i : clusters[c]) t += op_cost(ops[i]);
        return t;
    };
    // lane programs: per phase, per lane, list of op ids (-1 = NOP)
    std::vector<std::vector<std::vector<int>>> prog(n_ph, std::vector<std::vector<int>>(G));
    double useful = 0, total = 0;
    for (int ph = 0; ph < n_ph; ++ph) {
        std::map<std::vector<uint32_t>, std::vector<int>> groups;
        for (int c = 0; c < NC; ++c)
            if (lvl[c] == ph) groups[signature(c)].push_back(c);
        std::vector<std::pair<double, std::vector<int>>> full_rows;
        std::vector<int> leftovers;
        for (auto &kv : groups) {
            auto &v = kv.second;
            size_t i = 0;
            for (; i + G <= v.size(); i += G)
                full_rows.push_back({cl_cost(v[i]), std::vector<int>(v.begin() + i, v.begin() + i + G)});
            // A nearly full tail stays a row of its own (no divergence).
            if (v.size() - i >= (3 * G + 3) / 4) {
                full_rows.push_back({cl_cost(v[i]), std::vector<int>(v.begin() + i, v.end())});
            } else {
                for (; i < v.size(); ++i) leftovers.push_back(v[i]);
            }
        }
        // Heaviest leftovers first, packed into mixed rows.
        std::sort(leftovers.begin(), leftovers.end(), [&](int a, int b) { return cl_cost(a) > cl_cost(b); });
        for (size_t i = 0; i < leftovers.size(); i += G)
            full_rows.push_back({cl_cost(leftovers[i]),
                                 std::vector<int>(leftovers.begin() + i,
                                                  leftovers.begin() + std::min(leftovers.size(), i + (size_t)G))});
        std::stable_sort(full_rows.begin(), full_rows.end(),
                         [](const auto &a, const auto &b) { return a.first > b.first; });
        for (auto &row : full_rows) {
            size_t len = 0;
            for (int c : row.second) len = std::max(len, cl_len(c));
            double rc = 0;
            for (int c : row.second) rc = std::max(rc, cl_cost(c));
            for (uint32_t s = 0; s < G; ++s) {
                auto &lp = prog[ph][s];
                size_t k = 0;
                if (s < row.second.size()) {
                    const int c = row.second[s];
                    if (fused[c]) {
                        lp.push_back(-2 - c), ++k; // fused cluster marker
                    } else {
                        for (int i : clusters[c]) lp.push_back(i), ++k;
                    }
                    useful += cl_cost(c);
                }
                for (; k < len; ++k) lp.push_back(-1);
            }
            total += rc * G;
        }
    }

    // ---- 4b. device workspace layout ----
    // A state variable whose jet is only ever read at the CURRENT order (operands of
    // LINCOMB / ADDSUB / SVD / DIV numerators / fused pair differences) needs no
    // on-chip history: its jet goes to a global (L2-resident) scratch, written once
    // per order and read back once per step by the Horner update; on chip it keeps a
    // two-entry ping-pong (orders k, k+1).  Everything else is compacted.
    const uint32_t n_state = d.n_state;
    std::vector<char> hist(n_state, 0);
    auto mark_hist = [&](uint32_t ref) {
        if (ref == HY_REF_ONE) return;
        const uint32_t b = ref & 0x7fffffffu;
        if (b % P1 == 0 && b / P1 < n_state) hist[b / P1] = 1;
    };
    for (uint32_t i = 0; i < n_ops; ++i) {
        const hy_op &o = ops[i];
        switch (o.opcode) {
        case HY_OP_MUL: case HY_OP_INTG: mark_hist(o.a); mark_hist(o.b); break;
        case HY_OP_SQUARE: case HY_OP_POW: case HY_OP_SQRT: case HY_OP_EXP: case HY_OP_LOG: case HY_OP_SINCOS:
            mark_hist(o.a); break;
        case HY_OP_DIV: mark_hist(o.b); break;
        case HY_OP_SUMSQ: for (uint32_t j = 0; j < o.n; ++j) mark_hist(terms[o.b + j].src); break;
        case HY_OP_MULSH: mark_hist(o.a); for (uint32_t j = 0; j < o.n; ++j) mark_hist(terms[o.b + j].src); break;
        default: break;
        }
    }
    const bool allow_spill = std::getenv("HY_CUDA_NO_SPILL") == nullptr && spill_ok;
    // (event functions that ARE state variables need the full jet on chip)
    // collect every row block: base -> size
    std::map<uint32_t, uint32_t> blocks;
    auto add_block = [&](uint32_t ref, bool force_jet = false) {
        if (ref == HY_REF_ONE) return;
        const uint32_t b = ref & 0x7fffffffu;
        const uint32_t sz = ((ref & HY_REF_JET) || force_jet) ? P1 : 1u;
        auto it = blocks.find(b);
        if (it == blocks.end() || it->second < sz) blocks[b] = sz;
    };
    for (uint32_t i = 0; i < n_state; ++i) add_block((i * P1) | HY_REF_JET);
    for (uint32_t i = 0; i < n_ops; ++i) {
        const hy_op &o = ops[i];
        if (o.opcode != HY_OP_MULSH) add_block(o.dst);
        if (o.opcode == HY_OP_SINCOS) add_block(o.dst2);
        if (o.opcode == HY_OP_DIV || o.opcode == HY_OP_POW || o.opcode == HY_OP_SQRT || o.opcode == HY_OP_LOG)
            add_block(o.dst2 & 0x7fffffffu);
        if (is_term_op(o.opcode)) {
            for (uint32_t j = 0; j < o.n; ++j) {
                add_block(terms[o.b + j].src);
                if (o.opcode == HY_OP_MULSH) add_block(terms[o.b + j].dst);
            }
            if (o.opcode == HY_OP_MULSH) add_block(o.a);
        } else if (o.opcode != HY_OP_TIME) {
            add_block(o.a);
            if (o.opcode == HY_OP_MUL || o.opcode == HY_OP_DIV || o.opcode == HY_OP_ADDSUB || o.opcode == HY_OP_INTG)
                add_block(o.b);
        }
    }
    std::vector<char> is_ev_state(n_state, 0);
    for (uint32_t e = 0; e < d.n_events; ++e) {
        add_block(ev_ref_in[e], true);
        const uint32_t b = ev_ref_in[e] & 0x7fffffffu;
        if (b % P1 == 0 && b / P1 < n_state) is_ev_state[b / P1] = 1;
    }
    std::vector<int32_t> spill(n_state, -1);
    uint32_t n_spill = 0;
    for (uint32_t i = 0; i < n_state; ++i)
        if (allow_spill && !hist[i] && !is_ev_state[i]) spill[i] = (int32_t)n_spill++;
    std::map<uint32_t, uint32_t> new_base;
    uint32_t row_cnt = 0;
    for (auto &kv : blocks) {
        const uint32_t b = kv.first;
        uint32_t sz = kv.second;
        if (b % P1 == 0 && b / P1 < n_state && spill[b / P1] >= 0) sz = 2; // ping-pong
        new_base[b] = row_cnt;
        row_cnt += sz;
    }
    const uint32_t par_off = row_cnt, one_off = row_cnt + d.n_par;
    const uint32_t ws_len = one_off + P1;
    if (ws_len > 65535u) return "the system is too large: more than 65535 workspace rows per trajectory";
    // device reference: new base | JET | PP
    auto remap = [&](uint32_t ref) -> uint32_t {
        if (ref == HY_REF_ONE) return one_off | HY_DREF_JET;
        const uint32_t b = ref & 0x7fffffffu;
        auto it = new_base.find(b);
        const uint32_t nb = it == new_base.end() ? 0u : it->second;
        if (b % P1 == 0 && b / P1 < n_state && spill[b / P1] >= 0) return nb | HY_DREF_PP;
        return nb | ((ref & HY_REF_JET) ? HY_DREF_JET : 0u);
    };

    // ---- 5. emit ----
    out = Program();
    out.ws_len = ws_len;
    out.par_off = par_off;
    out.one_off = one_off;
    out.n_spill = n_spill;
    out.state_spill = spill;
    for (uint32_t i = 0; i < n_state; ++i) out.state_row.push_back(new_base[i * P1]);
    for (uint32_t e = 0; e < d.n_events; ++e) out.ev_ref.push_back(remap(ev_ref_in[e] | HY_REF_JET) & 0x3fffffffu);
    out.G = G;
    out.n_phases = (uint32_t)n_ph;
    out.n_clusters = (uint32_t)NC;
    out.lane_utilisation = total > 0 ? useful / total : 1.0;
    out.phase_slot.push_back(0);
    uint32_t slots = 0;
    for (int ph = 0; ph < n_ph; ++ph) {
        slots += (uint32_t)prog[ph][0].size();
        out.phase_slot.push_back(slots);
    }
    out.n_slots = slots;
    DOp nop{};
    nop.opcode = OP_NOP;
    out.ops.assign((size_t)slots * G, nop);
    std::vector<std::vector<DTerm>> lane_terms(G);
    std::map<uint64_t, uint16_t> imm_idx;
    auto imm_of = [&](double v) {
        uint64_t key;
        std::memcpy(&key, &v, 8);
        auto it = imm_idx.find(key);
        if (it != imm_idx.end()) return it->second;
        uint16_t id = (uint16_t)out.imm.size();
        out.imm.push_back(v);
        imm_idx[key] = id;
        return id;
    };
    auto off16 = [&](uint32_t dref) { return (uint16_t)(dref & 0x3fffffffu); };
    for (int ph = 0; ph < n_ph; ++ph)
        for (uint32_t s = 0; s < G; ++s) {
            const auto &lp = prog[ph][s];
            for (size_t j = 0; j < lp.size(); ++j) {
                if (lp[j] == -1) continue;
                if (lp[j] <= -2) {
                    // ---- fused pair-interaction cluster ----
                    const auto &m = clusters[-2 - lp[j]];
                    const size_t nd = m.size() - 3;
                    const hy_op &sq = ops[m[nd]], &pw = ops[m[nd + 1]], &ms = ops[m[nd + 2]];
                    DOp q{};
                    q.opcode = DOP_PAIR;
                    q.n = (uint16_t)nd;
                    q.a = off16(remap(sq.dst));
                    q.dst = off16(remap(pw.dst));
                    q.dst2 = off16(remap(pw.dst2 & 0x7fffffffu));
                    q.imm = imm_of(pw.imm);
                    q.b = (uint16_t)lane_terms[s].size();
                    for (size_t i = 0; i < nd; ++i) {
                        const hy_op &ad = ops[m[i]];
                        DTerm u{};
                        u.coef = (double)(((ad.flags & HY_OPF_NEGA) ? 1 : 0) | ((ad.flags & HY_OPF_NEGB) ? 2 : 0));
                        u.src = remap(ad.a);
                        u.aux = remap(ad.b);
                        lane_terms[s].push_back(u);
                        DTerm v{};
                        v.coef = 0;
                        v.src = remap(ad.dst);
                        v.aux = remap(terms[ms.b + i].dst);
                        lane_terms[s].push_back(v);
                    }
                    out.ops[((size_t)out.phase_slot[ph] + j) * G + s] = q;
                    continue;
                }
                const hy_op &o = ops[lp[j]];
                DOp q{};
                q.opcode = (uint8_t)o.opcode;
                q.flags = (uint8_t)(o.flags & 0xf);
                q.n = (uint16_t)o.n;
                auto setref = [&](uint32_t ref, uint16_t &field, uint8_t jflag, uint16_t pflag) {
                    const uint32_t r = remap(ref);
                    field = off16(r);
                    if (r & HY_DREF_JET) q.flags |= jflag;
                    if (r & HY_DREF_PP) q.pad |= pflag;
                };
                if (o.opcode != HY_OP_MULSH) setref(o.dst, q.dst, DF_JDST, DP_DST);
                const bool writes_state = o.opcode == HY_OP_SVD || (o.flags & HY_OPF_SVD);
                if (o.opcode == HY_OP_SINCOS)
                    setref(o.dst2, q.dst2, DF_JDST2, 0);
                else if (o.opcode == HY_OP_DIV || o.opcode == HY_OP_POW || o.opcode == HY_OP_SQRT || o.opcode == HY_OP_LOG)
                    q.dst2 = off16(remap(o.dst2 & 0x7fffffffu)); // scratch row holding 1/a[0]
                else if (writes_state) {
                    // spill slot + 1 of the destination state variable (0: resident)
                    const uint32_t sv = (o.dst & 0x7fffffffu) / P1;
                    q.dst2 = (uint16_t)(sv < n_state && spill[sv] >= 0 ? spill[sv] + 1 : 0);
                }
                if (o.opcode != HY_OP_TIME) setref(o.a, q.a, DF_JA, DP_A);
                if (is_term_op(o.opcode)) {
                    if (lane_terms[s].size() + o.n > 65535u) return "too many terms per lane";
                    q.b = (uint16_t)lane_terms[s].size();
                    bool has_par = false;
                    for (uint32_t t = 0; t < o.n; ++t) {
                        const hy_term &ht = terms[o.b + t];
                        if (ht.par >= (int32_t)d.n_par) return "parameter index out of bounds";
                        DTerm u{};
                        u.coef = ht.coef;
                        u.src = remap(ht.src);
                        if (o.opcode == HY_OP_MULSH) {
                            u.aux = remap(ht.dst);
                        } else if (o.opcode == HY_OP_LINCOMB) {
                            const uint32_t r = u.src;
                            const uint32_t mask = (r & HY_DREF_JET) ? 0xffffu : ((r & HY_DREF_PP) ? 1u : 0u);
                            u.src = r & 0x3fffffffu;
                            u.aux = (ht.par >= 0 ? par_off + (uint32_t)ht.par : one_off) | (mask << 16);
                            if (ht.par >= 0) has_par = true;
                        } else {
                            u.aux = ht.par >= 0 ? par_off + (uint32_t)ht.par : one_off;
                        }
                        lane_terms[s].push_back(u);
                    }
                    if (o.opcode == HY_OP_LINCOMB && !has_par) q.pad |= DP_NOPAR;
                } else {
                    setref(o.b, q.b, DF_JB, DP_B);
                }
                q.imm = imm_of(o.imm);
                (void)P1;
                out.ops[((size_t)out.phase_slot[ph] + j) * G + s] = q;
            }
        }
    size_t tmax = 0;
    for (auto &v : lane_terms) tmax = std::max(tmax, v.size());
    out.n_tslots = (uint32_t)tmax;
    DTerm zt{};
    zt.aux = one_off;
    out.terms.assign(tmax * G, zt);
    for (uint32_t s = 0; s < G; ++s)
        for (size_t c = 0; c < lane_terms[s].size(); ++c) out.terms[c * G + s] = lane_terms[s][c];
    if (out.imm.empty()) out.imm.push_back(0.0);
    return "";
}

} // namespace hy
