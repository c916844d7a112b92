// hy_evtape_host.hpp - host-side lowering of the event tape (include/hy_cuda.h, hy_event_tape)
// into the device program of hy_evtape.cuh: 32-byte ops, 16-byte terms, an immediate table and
// the per-event op ranges, in one blob that the kernel stages in shared memory.
//
// It also decides which ops must run at EVERY order (EOF_ALL): recurrences (div, pow, exp, ...)
// and every op whose history some other op reads.  A linear op / product / square whose history
// nobody reads runs at orders 0, p-1, p only - its output at order k is an explicit function of
// its operands' orders <= k, and those three orders are all a step needs unless an event may
// actually happen in it.  EOF_ALL ops never read the output of a three-order op, so the kernel
// runs them first, op by op over all orders, and then the three-order ops order by order.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/hy_cuda.h"
#include "hy_evtape.cuh"

namespace hy {

struct EvtProgram {
    std::vector<unsigned char> blob;
    uint32_t n_ops = 0, n_terms = 0, n_imm = 0, n_slots = 0;
    std::vector<uint32_t> ev_off; // [n_events] offset of every event jet inside the event workspace
    std::vector<uint32_t> state_used; // [n_state] 1: the event tape reads this state variable
    // the lowered program itself (for the code generator of hy_jit.hpp: events as straight-line code)
    std::vector<EOp> ops;
    std::vector<ETerm> terms;
    std::vector<double> imm;
    std::vector<uint32_t> op_start, ev_slot;
};

inline std::string build_event_program(uint32_t n_state, uint32_t order, const std::vector<hy_op> &ops,
                                       const std::vector<hy_term> &terms, const std::vector<uint32_t> &ev_ref,
                                       const std::vector<uint32_t> &op_start, uint32_t n_rows, EvtProgram &out)
{
    const uint32_t P1 = order + 1, base0 = n_state * P1;
    if (n_rows >= 0x3fffu) return "the event workspace is too large";
    auto eref = [&](uint32_t ref, bool &ok) -> uint16_t {
        if (ref == HY_REF_ONE) return ER_ONE;
        const uint32_t b = ref & 0x7fffffffu;
        if (b < base0) {
            if (b % P1 != 0 || !(ref & HY_REF_JET)) ok = false;
            return (uint16_t)(ER_STATE | (b / P1));
        }
        if (b - base0 >= n_rows) ok = false;
        return (uint16_t)(((ref & HY_REF_JET) ? ER_JET : ER_CUR) | (b - base0));
    };
    // interval slots: state variables first, then one per workspace block
    std::map<uint32_t, uint16_t> slot_of;
    uint32_t n_slots = n_state;
    auto slot = [&](uint32_t ref) -> uint16_t {
        if (ref == HY_REF_ONE) return 0;
        const uint32_t b = ref & 0x7fffffffu;
        if (b < base0) return (uint16_t)(b / P1);
        auto it = slot_of.find(b);
        if (it == slot_of.end()) it = slot_of.emplace(b, (uint16_t)n_slots++).first;
        return it->second;
    };
    const size_t n_ops = ops.size();
    std::vector<EOp> eo(n_ops);
    std::vector<ETerm> et(terms.size());
    std::vector<double> imm;
    bool ok = true;
    auto is_lin = [](uint16_t oc) { return oc == HY_OP_LINCOMB || oc == HY_OP_ADDSUB || oc == HY_OP_TIME; };
    auto is_term = [](uint16_t oc) { return oc == HY_OP_LINCOMB || oc == HY_OP_SUMSQ || oc == HY_OP_MULSH; };
    // producer of every workspace block, consumers of every op
    std::map<uint32_t, int> prod;
    for (size_t i = 0; i < n_ops; ++i) {
        const hy_op &o = ops[i];
        if (o.opcode >= HY_OP_COUNT || o.opcode == HY_OP_SVD) return "unsupported opcode in the event tape";
        if (is_term(o.opcode) && (uint64_t)o.b + o.n > terms.size()) return "term range out of bounds (event tape)";
        if (o.opcode == HY_OP_MULSH) {
            for (uint32_t j = 0; j < o.n; ++j) prod[terms[o.b + j].dst & 0x7fffffffu] = (int)i;
        } else {
            prod[o.dst & 0x7fffffffu] = (int)i;
            if (o.opcode == HY_OP_SINCOS) prod[o.dst2 & 0x7fffffffu] = (int)i;
        }
    }
    // need_all[i]: some consumer reads the history of op i's output (a non-linear consumer), or is a
    // linear op that itself is needed at every order.  Ops are in dependency order: walk backwards.
    std::vector<char> need_all(n_ops, 0);
    auto mark = [&](uint32_t ref, bool consumer_all) {
        if (ref == HY_REF_ONE) return;
        auto it = prod.find(ref & 0x7fffffffu);
        if (it != prod.end() && consumer_all) need_all[it->second] = 1;
    };
    for (size_t ii = n_ops; ii-- > 0;) {
        const hy_op &o = ops[ii];
        const bool c_all = !is_lin(o.opcode) || need_all[ii];
        if (is_term(o.opcode)) {
            for (uint32_t j = 0; j < o.n; ++j) mark(terms[o.b + j].src, c_all);
            if (o.opcode == HY_OP_MULSH) mark(o.a, c_all);
        } else if (o.opcode != HY_OP_TIME) {
            mark(o.a, c_all);
            if (o.opcode == HY_OP_MUL || o.opcode == HY_OP_DIV || o.opcode == HY_OP_ADDSUB) mark(o.b, c_all);
        }
    }
    for (size_t i = 0; i < n_ops; ++i) {
        const hy_op &o = ops[i];
        EOp q{};
        q.opcode = (uint8_t)o.opcode;
        q.flags = (uint8_t)(((o.flags & HY_OPF_NEGA) ? EOF_NEGA : 0) | ((o.flags & HY_OPF_NEGB) ? EOF_NEGB : 0));
        const bool explicit_op = o.opcode == HY_OP_MUL || o.opcode == HY_OP_SQUARE || o.opcode == HY_OP_SUMSQ ||
                                 o.opcode == HY_OP_MULSH;
        // recurrences (and TIME) run at every order; linear ops and explicit convolutions only if
        // somebody reads their history - otherwise orders 0, p-1, p are all a step needs
        const bool lin = is_lin(o.opcode) && o.opcode != HY_OP_TIME;
        if (!(explicit_op || lin) || need_all[i]) q.flags |= EOF_ALL;
        q.n = (uint16_t)o.n;
        if (o.opcode != HY_OP_MULSH) {
            q.dst = eref(o.dst, ok);
            q.sd = slot(o.dst);
            if ((q.dst & ER_KIND) == ER_STATE || (q.dst & ER_KIND) == ER_ONE) ok = false;
        }
        if (o.opcode == HY_OP_SINCOS) {
            q.dst2 = eref(o.dst2, ok);
            q.sd2 = slot(o.dst2);
        } else if (o.opcode == HY_OP_DIV || o.opcode == HY_OP_POW || o.opcode == HY_OP_SQRT || o.opcode == HY_OP_LOG) {
            q.dst2 = eref(o.dst2 & 0x7fffffffu, ok); // scratch row holding 1/a[0]
        }
        if (is_term(o.opcode)) {
            q.b = (uint16_t)o.b;
            for (uint32_t j = 0; j < o.n; ++j) {
                const hy_term &t = terms[o.b + j];
                if (t.par >= 0) return "runtime parameters in the event tape";
                ETerm u{};
                u.coef = t.coef;
                u.src = eref(t.src, ok);
                u.ssrc = slot(t.src);
                if (o.opcode == HY_OP_MULSH) {
                    u.dst = eref(t.dst, ok);
                    u.sdst = slot(t.dst);
                }
                et[o.b + j] = u;
            }
            if (o.opcode == HY_OP_MULSH) {
                q.a = eref(o.a, ok);
                q.sa = slot(o.a);
            }
        } else if (o.opcode != HY_OP_TIME) {
            q.a = eref(o.a, ok);
            q.sa = slot(o.a);
            if (o.opcode == HY_OP_MUL || o.opcode == HY_OP_DIV || o.opcode == HY_OP_ADDSUB) {
                q.b = eref(o.b, ok);
                q.sb = slot(o.b);
            }
        }
        if (o.opcode == HY_OP_POW) {
            q.imm = (uint16_t)imm.size();
            imm.push_back(o.imm);
        }
        eo[i] = q;
    }
    if (!ok) return "malformed row reference in the event tape";
    if (imm.empty()) imm.push_back(0.0);
    out = EvtProgram();
    out.n_ops = (uint32_t)n_ops;
    out.n_terms = (uint32_t)et.size();
    out.n_imm = (uint32_t)imm.size();
    out.n_slots = n_slots;
    const uint32_t n_ev = (uint32_t)ev_ref.size();
    for (uint32_t e = 0; e < n_ev; ++e) {
        const uint32_t b = ev_ref[e] & 0x7fffffffu;
        if (b < base0 || b - base0 + P1 > n_rows) return "an event jet lies outside the event workspace";
        out.ev_off.push_back(b - base0);
    }
    // state variables read by the event tape (only these need an enclosure over the step)
    out.state_used.assign(n_state, 0);
    for (const EOp &q : eo) {
        if ((q.a & ER_KIND) == ER_STATE && q.opcode != HY_OP_TIME && !is_term(q.opcode)) out.state_used[q.a & 0x3fff] = 1;
        if (q.opcode == HY_OP_MULSH && (q.a & ER_KIND) == ER_STATE) out.state_used[q.a & 0x3fff] = 1;
        if ((q.opcode == HY_OP_MUL || q.opcode == HY_OP_DIV || q.opcode == HY_OP_ADDSUB) && (q.b & ER_KIND) == ER_STATE)
            out.state_used[q.b & 0x3fff] = 1;
    }
    for (const ETerm &t : et)
        if ((t.src & ER_KIND) == ER_STATE) out.state_used[t.src & 0x3fff] = 1;
    out.ops = eo;
    out.terms = et;
    out.imm = imm;
    out.op_start = op_start;
    for (uint32_t e = 0; e < n_ev; ++e) out.ev_slot.push_back(slot(ev_ref[e]));
    // blob: [ops | terms | imm | op_start | ev_slot | state_used]
    const size_t bytes =
        n_ops * sizeof(EOp) + et.size() * sizeof(ETerm) + imm.size() * 8 + (n_ev + 1) * 4 + n_ev * 4 + n_state * 4;
    out.blob.assign((bytes + 15) / 16 * 16, 0);
    unsigned char *q = out.blob.data();
    std::memcpy(q, eo.data(), n_ops * sizeof(EOp));
    q += n_ops * sizeof(EOp);
    if (!et.empty()) std::memcpy(q, et.data(), et.size() * sizeof(ETerm));
    q += et.size() * sizeof(ETerm);
    std::memcpy(q, imm.data(), imm.size() * 8);
    q += imm.size() * 8;
    std::memcpy(q, op_start.data(), (n_ev + 1) * 4);
    q += (n_ev + 1) * 4;
    for (uint32_t e = 0; e < n_ev; ++e) {
        const uint32_t s = slot(ev_ref[e]);
        std::memcpy(q + 4 * e, &s, 4);
    }
    q += 4 * n_ev;
    for (uint32_t i = 0; i < n_state; ++i) {
        const uint32_t u = out.state_used[i];
        std::memcpy(q + 4 * i, &u, 4);
    }
    return "";
}

} // namespace hy
