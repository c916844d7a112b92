// hy_jit.hpp - run-time compiled kernels: the order sweep of ANY tape as straight-line CUDA.
//
// The reference JIT-compiles every ODE system to native code in the taylor_adaptive_batch
// constructor ([UPSTREAM] LLVM, called from /root/reference/heyoka/expose_batch_integrators.cpp:166-208).
// The B200 counterpart: hy_create lowers the scheduled program (hy_schedule.hpp, one lane per
// trajectory) to CUDA source - every op of the tape becomes one statement with literal row offsets,
// coefficients and term counts, the convolutions are the bodies of hy_kernels.cuh - and compiles it
// with NVRTC for sm_100a together with the persistent propagate kernel (hy_kernels.cuh, built with
// HY_JIT): the same step-size control, event detection, dense / continuous output and bookkeeping
// code as the precompiled kernels, with the tape interpreter replaced by the generated function.
//
// Execution model of the generated kernels: ONE THREAD PER TRAJECTORY.  The 32 trajectories of a
// warp execute the same instruction stream (no divergence between op kinds, no group
// synchronisation), their workspace is interleaved row by row (row r of lane l at [r * 32 + l]):
// every access of a warp is one contiguous 256-byte (FP64) segment - conflict-free in shared
// memory, fully coalesced in global memory.  Small systems keep the workspace in shared memory;
// systems whose jets do not fit stream them from global memory (L1/L2).
//
// Compiled kernels are cached on disk (cubin + lowered kernel name, keyed by a hash of the source
// and the options): <directory of libhy_cuda.so>/jit_cache, or $HY_CUDA_JIT_CACHE.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <numeric>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/hy_cuda.h"
#include "hy_schedule.hpp"
#include "hy_evtape_host.hpp"

namespace hy {
namespace jit {

constexpr uint32_t WS = 32; // element stride between rows of the warp-interleaved workspace

// --------------------------------------------------------------------------------------------
// Code generation
// --------------------------------------------------------------------------------------------
struct Gen {
    const hy_dims &d;
    const Program &pr;
    std::ostringstream os;
    // order blocking (blockconv_core, hy_kernels.cuh): block size M (0: off), first scratch row of
    // every blocked op (M - 1 rows per output), number of scratch rows
    uint32_t M = 0;
    uint32_t pf_dist = 0, pf_level = 1; // software prefetch: ops ahead (0: off), cache level
    uint32_t batch = 0;                  // loads per batch of independent linear ops (0: program order, no batches)
    bool div_group = true;               // quotients of a level by the same denominator as groups (jop_divsh)
    std::vector<int64_t> qbase;
    uint32_t q_rows = 0;
    Gen(const hy_dims &d_, const Program &p_, uint32_t blk) : d(d_), pr(p_)
    {
        qbase.assign(pr.n_slots, -1);
        if (blk >= 2 && d.order >= 2 * blk) {
            M = blk;
            uint32_t next = pr.ws_len;
            for (uint32_t i = 0; i < pr.n_slots; ++i) {
                const DOp &o = pr.ops[i];
                uint32_t outs = 0;
                if (o.opcode == HY_OP_MUL || o.opcode == HY_OP_DIV) outs = 1;
                if (o.opcode == HY_OP_MULSH && o.n >= 2 && o.n <= 4) outs = o.n;
                if (outs && next + outs * (M - 1) < 65535u) {
                    qbase[i] = next;
                    next += outs * (M - 1);
                }
            }
            q_rows = next - pr.ws_len;
        }
    }

    static std::string lit(double v)
    {
        char buf[64];
        std::snprintf(buf, sizeof buf, "%a", v);
        std::string s(buf);
        if (s.find("inf") != std::string::npos || s.find("nan") != std::string::npos) {
            // (coefficients of a tape are finite; keep the generator total anyway)
            return v != v ? "(0.0/0.0)" : (v > 0 ? "(1.0/0.0)" : "(-1.0/0.0)");
        }
        return "(R)" + s;
    }
    // row expression at order k: base (+ k for jets)
    static std::string row(uint32_t base, bool jet, const char *k = "k")
    {
        std::string s = std::to_string(base);
        if (jet) s += std::string(" + ") + k;
        return s;
    }
    static std::string Wr(uint32_t base, bool jet, const char *k = "k") { return "W(" + row(base, jet, k) + ")"; }
    static std::string ptr(uint32_t base) { return "(w + " + std::to_string(base) + " * HY_WS)"; }
    static std::string ptrk(uint32_t base, const char *k = "k")
    {
        return "(w + (" + std::to_string(base) + " + " + k + ") * HY_WS)";
    }

    void store(const DOp &o, const std::string &val)
    {
        if (o.opcode == HY_OP_SVD || (o.flags & HY_OPF_SVD))
            os << "        W(" << o.dst << " + k + 1) = (" << val << ") * rk[k + 1];\n";
        else
            os << "        " << Wr(o.dst, o.flags & DF_JDST) << " = " << val << ";\n";
    }

    // Prefetch the operand rows of op `o` at order k (plain sweep: the full history of every
    // convolution operand, the current row of the others).
    void emit_prefetch(const DOp &o, int lvl)
    {
        const DTerm *t = pr.terms.data() + o.b;
        const std::string L = std::to_string(lvl);
        auto jet = [&](uint32_t base) {
            os << "        pf_rows<R, HY_WS, " << L << ">(" << ptr(base) << ", (int)k + 1);\n";
        };
        auto cur = [&](uint32_t base, bool j) {
            os << "        pf_rows<R, HY_WS, " << L << ">(" << (j ? ptrk(base) : ptr(base)) << ", 1);\n";
        };
        switch (o.opcode) {
        case HY_OP_LINCOMB:
            for (uint32_t i = 0; i < o.n; ++i) cur(t[i].src, (t[i].aux >> 16) == 0xffffu);
            break;
        case HY_OP_ADDSUB: cur(o.a, o.flags & DF_JA); cur(o.b, o.flags & DF_JB); break;
        case HY_OP_SVD: cur(o.a, o.flags & DF_JA); break;
        case HY_OP_MUL: case HY_OP_INTG: jet(o.a); jet(o.b); break;
        case HY_OP_SQUARE: jet(o.a); break;
        case HY_OP_SUMSQ: for (uint32_t i = 0; i < o.n; ++i) jet(t[i].src & 0x3fffffffu); break;
        case HY_OP_MULSH: jet(o.a); for (uint32_t i = 0; i < o.n; ++i) jet(t[i].src & 0x3fffffffu); break;
        case HY_OP_DIV: jet(o.b); jet(o.dst); cur(o.a, o.flags & DF_JA); break;
        case HY_OP_POW: case HY_OP_SQRT: case HY_OP_EXP: case HY_OP_LOG: jet(o.a); jet(o.dst); break;
        case HY_OP_SINCOS: jet(o.a); jet(o.dst); jet(o.dst2); break;
        default: break;
        }
    }

    // ---- linear ops (LINCOMB / ADDSUB / SVD) in three phases, so that a batch of independent ops
    // can issue all its loads before the first FMA waits on one (one memory round trip per batch
    // instead of one per op).  Arithmetic as exec_op: one FMA chain in term order.
    static bool is_linear(const DOp &o) { return o.opcode == HY_OP_LINCOMB || o.opcode == HY_OP_ADDSUB || o.opcode == HY_OP_SVD; }
    uint32_t lin_nloads(const DOp &o) const
    {
        if (o.opcode == HY_OP_LINCOMB) {
            uint32_t n = o.n;
            const DTerm *t = pr.terms.data() + o.b;
            if (!(o.pad & DP_NOPAR))
                for (uint32_t i = 0; i < o.n; ++i) n += (t[i].aux & 0xffffu) != pr.one_off;
            return n;
        }
        return o.opcode == HY_OP_ADDSUB ? 2u : 1u;
    }
    void lin_loads(const DOp &o, uint32_t id)
    {
        const DTerm *t = pr.terms.data() + o.b;
        const std::string v = "v" + std::to_string(id) + "_";
        if (o.opcode == HY_OP_LINCOMB) {
            const bool nopar = (o.pad & DP_NOPAR) != 0;
            for (uint32_t u = 0; u < o.n; ++u) {
                const DTerm &tt = t[u];
                os << "        const R " << v << u << " = " << Wr(tt.src, (tt.aux >> 16) == 0xffffu) << ";\n";
                const uint32_t mult = tt.aux & 0xffffu;
                if (!nopar && mult != pr.one_off) os << "        const R " << v << "m" << u << " = W(" << mult << ");\n";
            }
        } else if (o.opcode == HY_OP_ADDSUB) {
            os << "        const R " << v << "a = " << ((o.flags & HY_OPF_NEGA) ? "-" : "") << Wr(o.a, o.flags & DF_JA) << ";\n";
            os << "        const R " << v << "b = " << ((o.flags & HY_OPF_NEGB) ? "-" : "") << Wr(o.b, o.flags & DF_JB) << ";\n";
        } else {
            os << "        const R " << v << "a = " << Wr(o.a, o.flags & DF_JA) << ";\n";
        }
    }
    void lin_math(const DOp &o, uint32_t id)
    {
        const DTerm *t = pr.terms.data() + o.b;
        const std::string v = "v" + std::to_string(id) + "_";
        if (o.opcode == HY_OP_LINCOMB) {
            const bool nopar = (o.pad & DP_NOPAR) != 0;
            os << "        R " << v << "acc = 0;\n";
            for (uint32_t u = 0; u < o.n; ++u) {
                const DTerm &tt = t[u];
                const uint32_t mult = tt.aux & 0xffffu;
                if (nopar || mult == pr.one_off)
                    os << "        " << v << "acc = r_fma(" << lit(tt.coef) << ", " << v << u << ", " << v << "acc);\n";
                else
                    os << "        " << v << "acc = r_fma(" << lit(tt.coef) << " * " << v << "m" << u << ", " << v << u << ", " << v
                       << "acc);\n";
            }
        } else if (o.opcode == HY_OP_ADDSUB) {
            os << "        const R " << v << "acc = " << v << "a + " << v << "b;\n";
        } else {
            os << "        const R " << v << "acc = " << v << "a;\n";
        }
    }
    void lin_store(const DOp &o, uint32_t id) { store(o, "v" + std::to_string(id) + "_acc"); }

    // ---- same-sweep dependencies: the rows an op reads / writes at the CURRENT order ----
    void cur_rows(const DOp &o, std::vector<uint32_t> &rd, std::vector<uint32_t> &wr) const
    {
        const DTerm *t = pr.terms.data() + o.b;
        const bool writes_next = o.opcode == HY_OP_SVD || (o.flags & HY_OPF_SVD); // x[k+1]: next sweep
        switch (o.opcode) {
        case HY_OP_LINCOMB:
            for (uint32_t i = 0; i < o.n; ++i) rd.push_back(t[i].src & 0x3fffffffu);
            break;
        case HY_OP_SUMSQ:
            for (uint32_t i = 0; i < o.n; ++i) rd.push_back(t[i].src & 0x3fffffffu);
            break;
        case HY_OP_MULSH:
            rd.push_back(o.a);
            for (uint32_t i = 0; i < o.n; ++i) {
                rd.push_back(t[i].src & 0x3fffffffu);
                wr.push_back(t[i].aux & 0x3fffffffu);
            }
            break;
        case HY_OP_ADDSUB: case HY_OP_MUL: case HY_OP_DIV: case HY_OP_INTG:
            rd.push_back(o.a);
            rd.push_back(o.b);
            break;
        case HY_OP_TIME: case OP_NOP: break;
        default: rd.push_back(o.a); break;
        }
        if (o.opcode != HY_OP_MULSH && o.opcode != OP_NOP && !writes_next) wr.push_back(o.dst);
        if (o.opcode == HY_OP_SINCOS) wr.push_back(o.dst2);
    }
    // Slots in dependency-level order (ops of one level are independent), linear ops first inside a
    // level, then grouped by kind.
    std::vector<uint32_t> level_order() const
    {
        std::map<uint32_t, uint32_t> prod; // row base -> slot
        std::vector<uint32_t> lvl(pr.n_slots, 0);
        for (uint32_t i = 0; i < pr.n_slots; ++i) {
            std::vector<uint32_t> rd, wr;
            cur_rows(pr.ops[i], rd, wr);
            uint32_t l = 0;
            for (uint32_t r : rd) {
                auto it = prod.find(r);
                if (it != prod.end()) l = std::max(l, lvl[it->second] + 1);
            }
            lvl[i] = l;
            for (uint32_t r : wr) prod[r] = i;
        }
        std::vector<uint32_t> idx(pr.n_slots);
        std::iota(idx.begin(), idx.end(), 0u);
        auto cls = [&](uint32_t i) { return is_linear(pr.ops[i]) ? 0u : 1u + pr.ops[i].opcode; };
        // (the quotients of a level next to each other by denominator: they are emitted as groups, jop_divsh)
        auto den = [&](uint32_t i) { return pr.ops[i].opcode == HY_OP_DIV ? pr.ops[i].b : 0u; };
        std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) {
            if (lvl[a] != lvl[b]) return lvl[a] < lvl[b];
            if (cls(a) != cls(b)) return cls(a) < cls(b);
            return den(a) < den(b);
        });
        levels_ = lvl;
        return idx;
    }
    mutable std::vector<uint32_t> levels_;

    // One op at order k (k: a variable of the generated code).  Linear ops are emitted inline (a few
    // loads and FMAs); everything with a convolution is a call to an out-of-line body (jop_*,
    // hy_kernels.cuh) with literal row numbers.  The unblocked arithmetic is exec_op's.
    bool emit_op(const DOp &o, uint32_t slot = 0, bool sweep = false)
    {
        const DTerm *t = pr.terms.data() + o.b; // (G = 1: the lane's stream is the term array)
        const bool blocked = sweep && M && qbase[slot] >= 0;
        const std::string Ms = std::to_string(blocked ? M : 1u);
        const std::string bi = blocked ? "bi" : "HY_NOBLK";
        const uint32_t qb = blocked ? (uint32_t)qbase[slot] : 0u;
        auto rowk = [&](uint32_t base, bool jet) { return row(base, jet); };
        os << "      { // op " << (int)o.opcode << "\n";
        switch (o.opcode) {
        case HY_OP_LINCOMB:
        case HY_OP_ADDSUB:
        case HY_OP_SVD:
            lin_loads(o, slot);
            lin_math(o, slot);
            lin_store(o, slot);
            break;
        case HY_OP_MUL:
            os << "        jop_mul<R, HY_WS, " << Ms << ">(w, k, " << bi << ", " << o.a << ", " << o.b << ", "
               << rowk(o.dst, o.flags & DF_JDST) << ", " << qb << ");\n";
            break;
        case HY_OP_SQUARE:
            os << "        jop_square<R, HY_WS>(w, k, " << o.a << ", " << rowk(o.dst, o.flags & DF_JDST) << ");\n";
            break;
        case HY_OP_SUMSQ:
            os << "        R acc = 0, acc2 = 0;\n";
            for (uint32_t i = 0; i < o.n; ++i)
                os << "        jop_sumsq_term<R, HY_WS>(w, k, " << (t[i].src & 0x3fffffffu) << ", acc, acc2);\n";
            store(o, "(acc + acc) + acc2");
            break;
        case HY_OP_MULSH: {
            // (the fusion pass builds groups of 2..4 products; anything else falls back to single products)
            if (o.n >= 1 && o.n <= 4) {
                os << "        const JRows<R, HY_WS, " << Ms << ", " << o.n << "> r = {{";
                for (uint32_t i = 0; i < o.n; ++i) os << (i ? ", " : "") << (t[i].src & 0x3fffffffu);
                os << "}, {";
                for (uint32_t i = 0; i < o.n; ++i) os << (i ? ", " : "") << rowk(t[i].aux & 0x3fffffffu, t[i].aux & HY_DREF_JET);
                os << "}};\n";
                os << "        jop_mulsh<R, HY_WS, " << Ms << ", " << o.n << ">(w, k, " << bi << ", " << o.a << ", r, " << qb << ");\n";
            } else {
                for (uint32_t i = 0; i < o.n; ++i)
                    os << "        jop_mul<R, HY_WS, 1>(w, k, HY_NOBLK, " << (t[i].src & 0x3fffffffu) << ", " << o.a << ", "
                       << rowk(t[i].aux & 0x3fffffffu, t[i].aux & HY_DREF_JET) << ", 0);\n";
            }
        } break;
        case HY_OP_DIV:
            os << "        jop_div<R, HY_WS, " << Ms << ">(w, k, " << bi << ", " << rowk(o.a, o.flags & DF_JA) << ", " << o.b << ", "
               << o.dst << ", " << o.dst2 << ", " << qb << ");\n";
            break;
        case HY_OP_POW:
        case HY_OP_SQRT: {
            const double alpha = o.opcode == HY_OP_SQRT ? 0.5 : pr.imm[o.imm];
            char ab[64];
            std::snprintf(ab, sizeof ab, "%a", alpha);
            os << "        jop_pow<R, HY_WS>(w, rk, k, " << o.a << ", " << o.dst << ", " << o.dst2 << ", " << ab << ", "
               << (o.opcode == HY_OP_SQRT ? 1 : 0) << ");\n";
        } break;
        case HY_OP_EXP: os << "        jop_exp<R, HY_WS>(w, rk, k, " << o.a << ", " << o.dst << ");\n"; break;
        case HY_OP_LOG:
            os << "        jop_log<R, HY_WS>(w, rk, k, " << o.a << ", " << o.dst << ", " << o.dst2 << ");\n";
            break;
        case HY_OP_SINCOS:
            os << "        jop_sincos<R, HY_WS>(w, rk, k, " << o.a << ", " << o.dst << ", " << o.dst2 << ");\n";
            break;
        case HY_OP_INTG:
            os << "        jop_intg<R, HY_WS>(w, rk, k, " << o.a << ", " << o.b << ", " << rowk(o.dst, o.flags & DF_JDST) << ", "
               << (int)pr.imm[o.imm] << ");\n";
            break;
        case HY_OP_TIME: os << "        W(" << o.dst << " + k) = k == 0 ? tm : (k == 1 ? (R)1 : (R)0);\n"; break;
        case OP_NOP: break;
        default: return false; // (fused pair ops are not generated: the program is built without fusion)
        }
        os << "      }\n";
        return true;
    }

    // The whole translation unit.  Returns "" if the program holds an op the generator does not know.
    std::string source(int fp_bits, bool smem, uint32_t threads)
    {
        os << "// generated by hy_jit.hpp: " << d.n_state << " state variables, order " << d.order << ", " << pr.n_slots
           << " ops, " << pr.ws_len << " workspace rows\n";
        os << "#define HY_JIT 1\n#define HY_WS " << WS << "\n#define HY_JIT_THREADS " << threads << "\n";
        if (!div_group) os << "#define HY_JIT_NOSHARE 1\n";
        os << "#include \"hy_kernels.cuh\"\n";
        os << "#define W(r) w[(r) * HY_WS]\n";
        os << "namespace hy {\n";
        os << "template <typename R> __device__ __forceinline__ void hy_gen_jets(R *__restrict__ w, const R *__restrict__ rk, const R tm)\n{\n";
        os << "    _Pragma(\"unroll 1\") for (uint32_t k = 0; k < " << d.order << "u; ++k) {\n";
        if (M) {
            // orders [M, floor(p / M) M) run in blocks of M; bi: position inside the block
            os << "      const uint32_t bi = (k >= " << M << "u && k < " << (d.order / M) * M << "u) ? k % " << M
               << "u : 0xffffffffu;\n";
        }
        if (!batch) {
            for (uint32_t i = 0; i < pr.n_slots; ++i) {
                if (pf_dist && i + pf_dist < pr.n_slots) emit_prefetch(pr.ops[i + pf_dist], pf_level);
                if (!emit_op(pr.ops[i], i, true)) return "";
            }
        } else {
            // dependency levels; the linear ops of a level in batches (loads, arithmetic, stores)
            const std::vector<uint32_t> ord = level_order();
            size_t i = 0;
            while (i < ord.size()) {
                const DOp &o = pr.ops[ord[i]];
                if (!is_linear(o)) {
                    // quotients of this level by the same denominator: groups of up to four share its loads
                    size_t g = i;
                    if (o.opcode == HY_OP_DIV && !M && div_group)
                        while (g < ord.size() && g - i < (threads > 256 ? 2u : 4u) && pr.ops[ord[g]].opcode == HY_OP_DIV && pr.ops[ord[g]].b == o.b &&
                               levels_[ord[g]] == levels_[ord[i]])
                            ++g;
                    if (g - i >= 2) {
                        const size_t nt = g - i;
                        os << "      { // " << nt << " quotients by row " << o.b << "\n        const JDivRows<" << nt << "> r = {{";
                        for (size_t q = i; q < g; ++q)
                            os << (q > i ? ", " : "") << row(pr.ops[ord[q]].a, pr.ops[ord[q]].flags & DF_JA);
                        os << "}, {";
                        for (size_t q = i; q < g; ++q) os << (q > i ? ", " : "") << pr.ops[ord[q]].dst;
                        os << "}, {";
                        for (size_t q = i; q < g; ++q) os << (q > i ? ", " : "") << pr.ops[ord[q]].dst2;
                        os << "}};\n        jop_divsh<R, HY_WS, " << nt << ">(w, k, " << o.b << ", r);\n      }\n";
                        i = g;
                        continue;
                    }
                    if (!emit_op(o, ord[i], true)) return "";
                    ++i;
                    continue;
                }
                size_t j = i;
                uint32_t loads = 0;
                while (j < ord.size() && is_linear(pr.ops[ord[j]]) && levels_[ord[j]] == levels_[ord[i]] &&
                       (j == i || loads + lin_nloads(pr.ops[ord[j]]) <= batch)) {
                    loads += lin_nloads(pr.ops[ord[j]]);
                    ++j;
                }
                os << "      { // " << (j - i) << " linear ops of level " << levels_[ord[i]] << "\n";
                for (size_t q = i; q < j; ++q) lin_loads(pr.ops[ord[q]], ord[q]);
                for (size_t q = i; q < j; ++q) lin_math(pr.ops[ord[q]], ord[q]);
                for (size_t q = i; q < j; ++q) lin_store(pr.ops[ord[q]], ord[q]);
                os << "      }\n";
                i = j;
            }
        }
        os << "    }\n}\n";
        os << "template <typename R> __device__ __forceinline__ void hy_gen_ev_sweep(R *__restrict__ w, const R *__restrict__ rk, const R tm)\n{\n";
        os << "    const uint32_t k = " << d.order << "u;\n    (void)k; (void)rk; (void)tm; (void)w;\n";
        if (d.n_events)
            for (uint32_t i = 0; i < pr.n_slots; ++i) {
                const DOp &o = pr.ops[i];
                if (!(o.flags & HY_OPF_EVENT) || (o.flags & HY_OPF_SVD) || o.opcode == HY_OP_SVD) continue;
                if (!emit_op(o)) return "";
            }
        os << "}\n";
        const char *R = fp_bits == 64 ? "double" : "float";
        os << "template __global__ void propagate_kernel<" << R << ", 1, " << (smem ? "true" : "false")
           << ", 0, false, NBR_PMAX, true>(const KParams<" << R << ">);\n";
        os << "} // namespace hy\n";
        return os.str();
    }
};

// --------------------------------------------------------------------------------------------
// Event functions of a register-resident kernel as generated code.  The event tape
// (hy_evtape_host.hpp) is evaluated by ONE lane per trajectory as straight-line code: literal ops
// (the interpreter's switch and reference decoding fold away), the convolutions of the three orders
// a step needs (0, p-1, p) fully unrolled with immediate offsets.
// --------------------------------------------------------------------------------------------
struct EvtGen {
    const EvtProgram &ep;
    const std::vector<uint32_t> &state_row;
    uint32_t p;
    std::ostringstream os;
    EvtGen(const EvtProgram &e, const std::vector<uint32_t> &sr, uint32_t order) : ep(e), state_row(sr), p(order) {}

    static std::string eop(const EOp &o)
    {
        std::ostringstream s;
        s << "EOp{" << (int)o.opcode << ", " << (int)o.flags << ", " << o.n << ", " << o.dst << ", " << o.dst2 << ", " << o.a
          << ", " << o.b << ", " << o.imm << ", " << o.sd << ", " << o.sd2 << ", " << o.sa << ", " << o.sb << ", {0, 0, 0, 0, 0}}";
        return s.str();
    }
    // pointer to order 0 of a reference and its stride between orders (as source text)
    std::string base(uint16_t ref, std::string &stride) const
    {
        const uint16_t kind = ref & ER_KIND, off = ref & 0x3fff;
        if (kind == ER_STATE) {
            stride = "XS";
            return "(C.w + " + std::to_string(state_row[off]) + ")";
        }
        stride = kind == ER_JET ? "1" : "0";
        return "(C.ews + " + std::to_string(off) + ")";
    }
    bool explicit_conv(const EOp &o) const
    {
        return o.opcode == HY_OP_MUL || o.opcode == HY_OP_SQUARE || o.opcode == HY_OP_SUMSQ || o.opcode == HY_OP_MULSH;
    }
    // a three-order op at the literal order K
    void emit_at(const EOp &o, uint32_t K)
    {
        const ETerm *t = ep.terms.data() + o.b;
        const std::string Ks = std::to_string(K);
        auto sq = [&](uint16_t ref, const std::string &acc, const std::string &acc2) {
            std::string st;
            const std::string pa = base(ref, st);
            const uint32_t half = (K + 1) >> 1;
            os << "      { const R *a = " << pa << ";\n";
            if (half) os << "        " << acc << " += evt_conv_ct<R, " << st << ", " << st << ", " << half << ">(a, a + " << Ks << " * " << st << ");\n";
            if ((K & 1u) == 0) os << "        { const R m = a[" << (K >> 1) << " * " << st << "]; " << acc2 << " = evt_fma(m, m, " << acc2 << "); }\n";
            os << "      }\n";
        };
        switch (o.opcode) {
        case HY_OP_MUL: {
            std::string sa, sb;
            const std::string pa = base(o.a, sa), pb = base(o.b, sb);
            os << "      C.st(" << o.dst << ", " << Ks << ", evt_conv_ct<R, " << sa << ", " << sb << ", " << K + 1 << ">(" << pa << ", "
               << pb << " + " << Ks << " * " << sb << "));\n";
        } break;
        case HY_OP_SQUARE:
            // evt_exec: acc = conv; acc = acc + acc; even k: acc = fma(m, m, acc)
            os << "      { R acc = 0;\n";
            {
                std::string st;
                const std::string pa = base(o.a, st);
                const uint32_t half = (K + 1) >> 1;
                if (half)
                    os << "        acc = evt_conv_ct<R, " << st << ", " << st << ", " << half << ">(" << pa << ", " << pa << " + " << Ks
                       << " * " << st << ");\n";
            }
            if ((K & 1u) == 0) {
                std::string st;
                const std::string pa = base(o.a, st);
                os << "        acc = acc + acc; { const R m = " << pa << "[" << (K >> 1) << " * " << st << "]; acc = evt_fma(m, m, acc); }\n";
            } else {
                os << "        acc = acc + acc;\n";
            }
            os << "        C.st(" << o.dst << ", " << Ks << ", acc); }\n";
            break;
        case HY_OP_SUMSQ:
            os << "      { R acc = 0, acc2 = 0;\n";
            for (uint32_t i = 0; i < o.n; ++i) sq(t[i].src, "acc", "acc2");
            os << "        C.st(" << o.dst << ", " << Ks << ", (acc + acc) + acc2); }\n";
            break;
        case HY_OP_MULSH: {
            std::string sb;
            const std::string pb = base(o.a, sb);
            for (uint32_t i = 0; i < o.n; ++i) {
                std::string sa;
                const std::string pa = base(t[i].src, sa);
                os << "      C.st(" << t[i].dst << ", " << Ks << ", evt_conv_ct<R, " << sa << ", " << sb << ", " << K + 1 << ">(" << pa
                   << ", " << pb << " + " << Ks << " * " << sb << "));\n";
            }
        } break;
        default: os << "      evt_exec<R, XS>(" << eop(o) << ", terms, C, " << Ks << "u);\n"; break;
        }
    }
    // ------------------------------------------------------------------------------------------
    // Plan of the lane-parallel form.  Three-order squares / products ("units") read only state jets and
    // jets of every-order ops, never each other; they are de-duplicated (events often share y^2, z^2, ...),
    // spread over the lanes of the trajectory's group and computed from ONE pass over the operands'
    // coefficients; a LINCOMB "1 * x + c" that only units read is folded into their operand (alias).
    // The three orders of every unit go to a scratch area (the interval scratch, unused at that
    // point); lane 0 then runs the remaining three-order ops (LINCOMB / ADDSUB as literal code).
    // ------------------------------------------------------------------------------------------
    struct Opnd {
        bool state = false; // C.w + off, stride XS;  else C.ews + off, stride 1
        uint32_t off = 0;
        bool al = false;
        double c = 0;
        bool operator<(const Opnd &o) const
        {
            uint64_t a, b;
            std::memcpy(&a, &c, 8);
            std::memcpy(&b, &o.c, 8);
            return std::tie(state, off, al, a) < std::tie(o.state, o.off, o.al, b);
        }
    };
    struct Unit {
        int kind; // 0: square, 1: product
        Opnd a, b;
    };
    uint32_t group = 2, scratch_len = 0;
    std::vector<Unit> units;            // unique units
    std::map<uint32_t, uint32_t> unit_of; // ews offset of a unit op's output -> unit
    std::vector<char> alias;            // per op: folded into its consumers
    std::vector<Opnd> alias_op;         // per op (alias): the operand it stands for
    bool plan_ok = false;

    static std::string lit(double v)
    {
        char buf[64];
        std::snprintf(buf, sizeof buf, "%a", v);
        return std::string("(R)") + buf;
    }
    static bool is_ews(uint16_t ref) { return (ref & ER_KIND) == ER_CUR || (ref & ER_KIND) == ER_JET; }
    bool plan()
    {
        const size_t n_ops = ep.ops.size();
        alias.assign(n_ops, 0);
        alias_op.assign(n_ops, Opnd());
        std::map<uint32_t, size_t> prod; // ews offset -> producing op
        for (size_t i = 0; i < n_ops; ++i) {
            const EOp &o = ep.ops[i];
            if (o.opcode == HY_OP_MULSH) {
                for (uint32_t j = 0; j < o.n; ++j) prod[ep.terms[o.b + j].dst & 0x3fff] = i;
            } else {
                prod[o.dst & 0x3fff] = i;
                if (o.opcode == HY_OP_SINCOS) prod[o.dst2 & 0x3fff] = i;
            }
        }
        auto is_term = [](uint8_t oc) { return oc == HY_OP_LINCOMB || oc == HY_OP_SUMSQ || oc == HY_OP_MULSH; };
        auto unit_op = [](const EOp &o) {
            return !(o.flags & EOF_ALL) && (o.opcode == HY_OP_SQUARE || o.opcode == HY_OP_MUL || o.opcode == HY_OP_MULSH);
        };
        // alias candidates
        for (size_t i = 0; i < n_ops; ++i) {
            const EOp &o = ep.ops[i];
            if (!(o.flags & EOF_ALL) || o.opcode != HY_OP_LINCOMB || (o.dst & ER_KIND) != ER_JET) continue;
            int n_var = 0, n_one = 0;
            Opnd q;
            bool ok = true;
            for (uint32_t j = 0; j < o.n; ++j) {
                const ETerm &t = ep.terms[o.b + j];
                if ((t.src & ER_KIND) == ER_ONE) {
                    ++n_one;
                    q.c = t.coef;
                } else if ((t.src & ER_KIND) == ER_STATE && t.coef == 1.0) {
                    ++n_var;
                    q.state = true;
                    q.off = state_row[t.src & 0x3fff];
                } else
                    ok = false;
            }
            if (!ok || n_var != 1 || n_one > 1) continue;
            q.al = n_one == 1;
            // every reader must be a unit
            bool only_units = true;
            const uint16_t me = o.dst & 0x3fff;
            auto reads = [&](uint16_t ref) { return is_ews(ref) && (ref & 0x3fff) == me; };
            for (size_t j = 0; j < n_ops && only_units; ++j) {
                const EOp &c = ep.ops[j];
                bool r = false;
                if (is_term(c.opcode)) {
                    for (uint32_t u = 0; u < c.n; ++u) r = r || reads(ep.terms[c.b + u].src);
                    if (c.opcode == HY_OP_MULSH) r = r || reads(c.a);
                } else if (c.opcode != HY_OP_TIME) {
                    r = reads(c.a);
                    if (c.opcode == HY_OP_MUL || c.opcode == HY_OP_DIV || c.opcode == HY_OP_ADDSUB) r = r || reads(c.b);
                }
                if (r && !unit_op(c)) only_units = false;
            }
            for (uint32_t off : ep.ev_off)
                if (off == me) only_units = false; // (an event function itself)
            if (!only_units) continue;
            alias[i] = 1;
            alias_op[i] = q;
        }
        // units
        bool ok = true;
        auto opnd = [&](uint16_t ref) -> Opnd {
            Opnd q;
            const uint16_t kind = ref & ER_KIND, off = ref & 0x3fff;
            if (kind == ER_STATE) {
                q.state = true;
                q.off = state_row[off];
            } else if (kind == ER_JET) {
                auto it = prod.find(off);
                if (it != prod.end() && alias[it->second])
                    q = alias_op[it->second];
                else
                    q.off = off;
            } else
                ok = false; // (a product reading a single row or the unit jet: the serial form handles it)
            return q;
        };
        std::map<std::tuple<int, Opnd, Opnd>, uint32_t> seen;
        auto add = [&](int kind, Opnd a, Opnd b, uint16_t dst) {
            auto key = std::make_tuple(kind, a, b);
            auto it = seen.find(key);
            if (it == seen.end()) {
                it = seen.emplace(key, (uint32_t)units.size()).first;
                units.push_back(Unit{kind, a, b});
            }
            unit_of[dst & 0x3fff] = it->second;
        };
        for (size_t i = 0; i < n_ops; ++i) {
            const EOp &o = ep.ops[i];
            if (!unit_op(o)) continue;
            if (o.opcode == HY_OP_SQUARE)
                add(0, opnd(o.a), Opnd(), o.dst);
            else if (o.opcode == HY_OP_MUL)
                add(1, opnd(o.a), opnd(o.b), o.dst);
            else
                for (uint32_t j = 0; j < o.n; ++j) add(1, opnd(ep.terms[o.b + j].src), opnd(o.a), ep.terms[o.b + j].dst);
        }
        // a reader of a unit's output other than LINCOMB / ADDSUB (none is generated by the front end)
        for (size_t i = 0; i < n_ops && ok; ++i) {
            const EOp &o = ep.ops[i];
            if (o.opcode == HY_OP_LINCOMB || o.opcode == HY_OP_ADDSUB || o.opcode == HY_OP_TIME) continue;
            auto bad = [&](uint16_t ref) { return is_ews(ref) && unit_of.count(ref & 0x3fff); };
            if (is_term(o.opcode)) {
                for (uint32_t u = 0; u < o.n; ++u)
                    if (bad(ep.terms[o.b + u].src)) ok = false;
                if (o.opcode == HY_OP_MULSH && bad(o.a)) ok = false;
            } else {
                if (bad(o.a)) ok = false;
                if ((o.opcode == HY_OP_MUL || o.opcode == HY_OP_DIV) && bad(o.b)) ok = false;
            }
        }
        if (p < 2 || units.empty() || 3u * units.size() > scratch_len) ok = false;
        plan_ok = ok;
        if (!ok) {
            alias.assign(n_ops, 0);
            units.clear();
            unit_of.clear();
        }
        return ok;
    }
    // "sub == 0 ? v0 : sub == 1 ? v1 : ... : v_last"
    static std::string sel(const std::vector<std::string> &v)
    {
        bool same = true;
        for (const auto &x : v) same = same && x == v[0];
        if (same) return v[0];
        std::string s;
        for (size_t i = 0; i + 1 < v.size(); ++i) s += "sub == " + std::to_string(i) + "u ? " + v[i] + " : ";
        return s + v.back();
    }
    // value of an operand of a three-order LINCOMB / ADDSUB at the literal order K (q = 0, 1, 2)
    std::string val(uint16_t ref, uint32_t K, uint32_t q) const
    {
        const uint16_t kind = ref & ER_KIND, off = ref & 0x3fff;
        if (kind == ER_ONE) return K == 0 ? "(R)1" : "(R)0";
        if (kind == ER_STATE) return "C.w[" + std::to_string(state_row[off] ) + " + " + std::to_string(K) + " * XS]";
        auto it = unit_of.find(off);
        if (it != unit_of.end()) return "scr[" + std::to_string(3 * it->second + q) + "]";
        return "C.ews[" + std::to_string(off + (kind == ER_JET ? K : 0u)) + "]";
    }
    void emit_units()
    {
        const uint32_t NL = std::min<uint32_t>(group, 8u);
        // group by signature: same code on every lane of a round
        std::map<std::tuple<int, bool, bool>, std::vector<uint32_t>> sig;
        for (uint32_t u = 0; u < units.size(); ++u) sig[std::make_tuple(units[u].kind, units[u].a.state, units[u].b.state)].push_back(u);
        for (const auto &g : sig) {
            const int kind = std::get<0>(g.first);
            const std::vector<uint32_t> &us = g.second;
            for (size_t r0 = 0; r0 < us.size(); r0 += NL) {
                const size_t cnt = std::min<size_t>(NL, us.size() - r0);
                std::vector<std::string> oa, ala, ca, ob, alb, cb, so;
                bool any_ala = false, any_alb = false;
                for (size_t l = 0; l < cnt; ++l) {
                    const Unit &U = units[us[r0 + l]];
                    oa.push_back(std::to_string(U.a.off) + "u");
                    ala.push_back(U.a.al ? "true" : "false");
                    ca.push_back(lit(U.a.al ? U.a.c : 0.0));
                    ob.push_back(std::to_string(U.b.off) + "u");
                    alb.push_back(U.b.al ? "true" : "false");
                    cb.push_back(lit(U.b.al ? U.b.c : 0.0));
                    so.push_back(std::to_string(3 * us[r0 + l]) + "u");
                    any_ala = any_ala || U.a.al;
                    any_alb = any_alb || U.b.al;
                }
                const Unit &U0 = units[us[r0]];
                const std::string sa = U0.a.state ? "XS" : "1", sb = U0.b.state ? "XS" : "1";
                const std::string ba = U0.a.state ? "C.w" : "C.ews", bb = U0.b.state ? "C.w" : "C.ews";
                os << "    { // " << cnt << (kind ? " product(s)" : " square(s)") << "\n";
                os << "      const bool on = sub < " << cnt << "u;\n";
                os << "      const R *a = " << ba << " + (" << sel(oa) << ");\n";
                if (any_ala) os << "      const bool ala = " << sel(ala) << "; const R ca = " << sel(ca) << ";\n";
                if (kind) {
                    os << "      const R *b = " << bb << " + (" << sel(ob) << ");\n";
                    if (any_alb) os << "      const bool alb = " << sel(alb) << "; const R cb = " << sel(cb) << ";\n";
                }
                os << "      R *out = scr + (" << sel(so) << ");\n";
                if (kind)
                    os << "      evt_unit_mul<R, " << sa << ", " << sb << ", " << p << ">(a, " << (any_ala ? "ala, ca" : "false, (R)0")
                       << ", b, " << (any_alb ? "alb, cb" : "false, (R)0") << ", out, on);\n";
                else
                    os << "      evt_unit_sq<R, " << sa << ", " << p << ">(a, " << (any_ala ? "ala, ca" : "false, (R)0") << ", out, on);\n";
                os << "    }\n";
            }
        }
    }
    // lane 0: the three-order ops that are not units, at the literal order K
    void emit_rest_at(const EOp &o, uint32_t K, uint32_t q)
    {
        const ETerm *t = ep.terms.data() + o.b;
        switch (o.opcode) {
        case HY_OP_SQUARE:
        case HY_OP_MUL:
            if ((o.dst & ER_KIND) == ER_JET) os << "      C.st(" << o.dst << ", " << K << ", " << val(o.dst, K, q) << ");\n";
            break;
        case HY_OP_MULSH:
            for (uint32_t j = 0; j < o.n; ++j)
                if ((t[j].dst & ER_KIND) == ER_JET) os << "      C.st(" << t[j].dst << ", " << K << ", " << val(t[j].dst, K, q) << ");\n";
            break;
        case HY_OP_LINCOMB:
            os << "      { R acc = 0;\n";
            for (uint32_t j = 0; j < o.n; ++j) os << "        acc = evt_fma(" << lit(t[j].coef) << ", " << val(t[j].src, K, q) << ", acc);\n";
            os << "        C.st(" << o.dst << ", " << K << ", acc); }\n";
            break;
        case HY_OP_ADDSUB:
            os << "      C.st(" << o.dst << ", " << K << ", " << ((o.flags & EOF_NEGA) ? "-" : "") << val(o.a, K, q) << " + "
               << ((o.flags & EOF_NEGB) ? "-" : "") << val(o.b, K, q) << ");\n";
            break;
        default: emit_at(o, K); break;
        }
    }
    std::string source()
    {
        plan();
        const uint32_t P3[3] = {0u, p - 1u, p};
        os << "// event functions: " << ep.ops.size() << " ops, " << ep.ev_slot.size() << " events, order " << p << "; "
           << (plan_ok ? std::to_string(units.size()) + " product units on the lanes of the group" : std::string("serial form")) << "\n";
        os << "template <typename R, int XS>\n__device__ __forceinline__ void hy_gen_evt_norms(const EvtCtx<R, XS> &C, const ETerm "
              "*terms, R *scr, uint32_t sub)\n{\n    (void)terms; (void)scr;\n";
        if (plan_ok) {
            bool any_all = false;
            for (size_t i = 0; i < ep.ops.size(); ++i) any_all = any_all || ((ep.ops[i].flags & EOF_ALL) && !alias[i]);
            if (any_all) {
                os << "    if (sub == 0) {\n";
                for (size_t i = 0; i < ep.ops.size(); ++i)
                    if ((ep.ops[i].flags & EOF_ALL) && !alias[i])
                        os << "      evt_exec_all<R, XS>(" << eop(ep.ops[i]) << ", terms, C, " << p << "u);\n";
                os << "    }\n    __syncwarp();\n";
            }
            emit_units();
            os << "    __syncwarp();\n    if (sub == 0) {\n";
            for (uint32_t q = 0; q < 3; ++q) {
                os << "    { // order " << P3[q] << "\n";
                for (const EOp &o : ep.ops)
                    if (!(o.flags & EOF_ALL)) emit_rest_at(o, P3[q], q);
                os << "    }\n";
            }
            os << "    }\n";
        } else {
            os << "    if (sub != 0) return;\n";
            for (const EOp &o : ep.ops)
                if (o.flags & EOF_ALL) os << "    evt_exec_all<R, XS>(" << eop(o) << ", terms, C, " << p << "u);\n";
            for (uint32_t K : P3) {
                os << "    { // order " << K << "\n";
                for (const EOp &o : ep.ops)
                    if (!(o.flags & EOF_ALL)) emit_at(o, K);
                os << "    }\n";
            }
        }
        os << "}\n";
        // the remaining orders (rare: an event may happen in this step), lane 0
        os << "template <typename R, int XS>\n__device__ __forceinline__ void hy_gen_evt_order(const EvtCtx<R, XS> &C, const ETerm *terms, uint32_t k)\n{\n";
        os << "    (void)C; (void)terms; (void)k;\n";
        bool any_alias = false;
        for (size_t i = 0; i < ep.ops.size(); ++i) any_alias = any_alias || alias[i];
        if (any_alias) {
            os << "    if (k == 1u) { // (the folded operands as jets: the generic ops below read them)\n";
            for (size_t i = 0; i < ep.ops.size(); ++i)
                if (alias[i]) os << "      evt_exec_all<R, XS>(" << eop(ep.ops[i]) << ", terms, C, " << p << "u);\n";
            os << "    }\n";
        }
        for (const EOp &o : ep.ops)
            if (!(o.flags & EOF_ALL)) os << "    evt_exec<R, XS>(" << eop(o) << ", terms, C, k);\n";
        os << "}\n";
        emit_interval();
        return os.str();
    }
    // Interval pass: the state enclosures on the lanes of the group, then the tape on lane 0 - intervals in
    // registers and literal coefficients when every op is algebraic, the generic evt_interval otherwise.
    void emit_interval()
    {
        os << "template <typename R, int XS>\n__device__ __forceinline__ bool hy_gen_evt_interval(const R *w, R *iv, const ETerm *terms, "
              "const double *imm, R t0, R h, uint32_t sub)\n{\n    (void)terms; (void)imm; (void)t0;\n";
        std::vector<uint32_t> used;
        for (size_t i = 0; i < ep.state_used.size(); ++i)
            if (ep.state_used[i]) used.push_back((uint32_t)i);
        const uint32_t NL = std::min<uint32_t>(group, 8u);
        for (size_t r0 = 0; r0 < used.size(); r0 += NL) {
            const size_t cnt = std::min<size_t>(NL, used.size() - r0);
            std::vector<std::string> row, slot;
            for (size_t l = 0; l < cnt; ++l) {
                row.push_back(std::to_string(state_row[used[r0 + l]]) + "u");
                slot.push_back(std::to_string(2 * used[r0 + l]) + "u");
            }
            os << "    { const Ival<R> v = iv_taylor_abs<R, XS, " << p << ">(w + (" << sel(row) << "), h);\n";
            os << "      if (sub < " << cnt << "u) { R *o = iv + (" << sel(slot) << "); o[0] = v.lo; o[1] = v.hi; } }\n";
        }
        os << "    __syncwarp();\n    bool maybe = false;\n    if (sub == 0) {\n";
        bool simple = true;
        for (const EOp &o : ep.ops)
            simple = simple && (o.opcode == HY_OP_LINCOMB || o.opcode == HY_OP_ADDSUB || o.opcode == HY_OP_MUL ||
                                o.opcode == HY_OP_SQUARE || o.opcode == HY_OP_SUMSQ || o.opcode == HY_OP_MULSH);
        if (simple) {
            std::map<uint32_t, std::string> name; // slot -> variable
            for (uint32_t i : used) {
                name[i] = "s" + std::to_string(i);
                os << "      const Ival<R> s" << i << " = {iv[" << 2 * i << "], iv[" << 2 * i + 1 << "]};\n";
            }
            std::map<std::tuple<int, std::string, std::string>, std::string> seen;
            auto get = [&](uint16_t ref, uint16_t slot) -> std::string {
                if ((ref & ER_KIND) == ER_ONE) return "Ival<R>{(R)1, (R)1}";
                auto it = name.find(slot);
                return it == name.end() ? std::string("iv_all<R>()") : it->second;
            };
            auto def = [&](uint16_t slot, int kind, const std::string &a, const std::string &b, const std::string &expr) {
                auto key = std::make_tuple(kind, a, b);
                if (kind >= 0) {
                    auto it = seen.find(key);
                    if (it != seen.end()) {
                        name[slot] = it->second;
                        return;
                    }
                }
                const std::string nm = "s" + std::to_string(slot);
                os << "      const Ival<R> " << nm << " = iv_fix<R>(" << expr << ");\n";
                name[slot] = nm;
                if (kind >= 0) seen[key] = nm;
            };
            for (const EOp &o : ep.ops) {
                const ETerm *t = ep.terms.data() + o.b;
                switch (o.opcode) {
                case HY_OP_SQUARE: {
                    const std::string a = get(o.a, o.sa);
                    def(o.sd, 0, a, "", "iv_sqr<R>(" + a + ")");
                } break;
                case HY_OP_MUL: {
                    const std::string a = get(o.a, o.sa), b = get(o.b, o.sb);
                    def(o.sd, 1, a, b, "iv_mul<R>(" + a + ", " + b + ")");
                } break;
                case HY_OP_MULSH: {
                    const std::string b = get(o.a, o.sa);
                    for (uint32_t j = 0; j < o.n; ++j) {
                        const std::string a = get(t[j].src, t[j].ssrc);
                        def(t[j].sdst, 1, a, b, "iv_mul<R>(" + a + ", " + b + ")");
                    }
                } break;
                case HY_OP_ADDSUB: {
                    const std::string a = get(o.a, o.sa), b = get(o.b, o.sb);
                    const std::string alo = (o.flags & EOF_NEGA) ? "-" + a + ".hi" : a + ".lo", ahi = (o.flags & EOF_NEGA) ? "-" + a + ".lo" : a + ".hi";
                    const std::string blo = (o.flags & EOF_NEGB) ? "-" + b + ".hi" : b + ".lo", bhi = (o.flags & EOF_NEGB) ? "-" + b + ".lo" : b + ".hi";
                    def(o.sd, -1, "", "", "iv_widen<R>(" + alo + " + " + blo + ", " + ahi + " + " + bhi + ")");
                } break;
                case HY_OP_SUMSQ: {
                    os << "      R l" << o.sd << " = 0, h" << o.sd << " = 0;\n";
                    for (uint32_t j = 0; j < o.n; ++j)
                        os << "      { const Ival<R> q = iv_sqr<R>(" << get(t[j].src, t[j].ssrc) << "); l" << o.sd << " += q.lo; h" << o.sd
                           << " += q.hi; }\n";
                    def(o.sd, -1, "", "", "iv_widen<R>(l" + std::to_string(o.sd) + ", h" + std::to_string(o.sd) + ")");
                } break;
                default: { // LINCOMB
                    os << "      R l" << o.sd << " = 0, h" << o.sd << " = 0;\n";
                    for (uint32_t j = 0; j < o.n; ++j)
                        os << "      iv_lin_term<R>(l" << o.sd << ", h" << o.sd << ", " << lit(t[j].coef) << ", " << get(t[j].src, t[j].ssrc)
                           << ");\n";
                    os << "      const Ival<R> s" << o.sd << " = iv_lin_finish<R>(l" << o.sd << ", h" << o.sd << ", " << o.n << ");\n";
                    name[o.sd] = "s" + std::to_string(o.sd);
                } break;
                }
            }
            for (uint32_t sl : ep.ev_slot) {
                const std::string g = name.count(sl) ? name[sl] : std::string("iv_all<R>()");
                os << "      { const Ival<R> g = " << g << "; if (!(g.lo > (R)0 || g.hi < (R)0)) maybe = true; }\n";
            }
        } else {
            for (const EOp &o : ep.ops) os << "      evt_interval<R>(" << eop(o) << ", terms, iv, imm, t0, h);\n";
            for (uint32_t sl : ep.ev_slot)
                os << "      { const R glo = iv[" << 2 * sl << "], ghi = iv[" << 2 * sl + 1
                   << "]; if (!(glo > (R)0 || ghi < (R)0)) maybe = true; }\n";
        }
        os << "    }\n    return maybe;\n}\n";
    }
};

// Translation unit of a register-resident kernel (FX build) with generated event functions.
// `group`: lanes per trajectory of the kernel; `scratch_len`: elements of the interval scratch (it doubles as
// the exchange area of the lane-parallel products).
inline std::string evt_kernel_source(const EvtProgram &ep, const std::vector<uint32_t> &state_row, uint32_t order,
                                     const std::string &defs = "", uint32_t group = 1, uint32_t scratch_len = 0)
{
    EvtGen g(ep, state_row, order);
    g.group = group;
    g.scratch_len = scratch_len;
    std::string s = defs + "#define HY_JIT_EVT 1\n#include \"hy_kernels.cuh\"\nnamespace hy {\n";
    s += g.source();
    s += "} // namespace hy\n";
    return s;
}

inline std::string kernel_name(int fp_bits, bool smem)
{
    return std::string("hy::propagate_kernel<") + (fp_bits == 64 ? "double" : "float") + ", 1, " + (smem ? "true" : "false") +
           ", 0, false, hy::NBR_PMAX, true>";
}

// --------------------------------------------------------------------------------------------
// NVRTC (loaded on first use: libhy_cuda.so itself does not depend on it)
// --------------------------------------------------------------------------------------------
struct Nvrtc {
    void *h = nullptr;
    int (*CreateProgram)(void **, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    int (*DestroyProgram)(void **) = nullptr;
    int (*CompileProgram)(void *, int, const char *const *) = nullptr;
    int (*GetProgramLogSize)(void *, size_t *) = nullptr;
    int (*GetProgramLog)(void *, char *) = nullptr;
    int (*GetCUBINSize)(void *, size_t *) = nullptr;
    int (*GetCUBIN)(void *, char *) = nullptr;
    int (*AddNameExpression)(void *, const char *) = nullptr;
    int (*GetLoweredName)(void *, const char *, const char **) = nullptr;
    int (*Version)(int *, int *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string err;
    std::mutex mtx; // (contexts are created from several host threads: ensembles, the lane-sharded integrator)
    bool loaded = false;
    bool load()
    {
        std::lock_guard<std::mutex> lk(mtx);
        if (loaded) return true;
        if (h) return false; // (an earlier attempt found the library but not every symbol)
        if (const char *e = std::getenv("HY_CUDA_NVRTC_LIB")) { // (explicit library; also how the tests take NVRTC away)
            h = dlopen(e, RTLD_NOW | RTLD_LOCAL);
        } else {
            for (const char *nm : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
                h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
                if (h) break;
            }
        }
        if (!h) {
            err = "libnvrtc not found";
            return false;
        }
#define HY_NVRTC_SYM(f)                                                        \
    f = reinterpret_cast<decltype(f)>(dlsym(h, "nvrtc" #f));                   \
    if (!f) {                                                                  \
        err = "nvrtc" #f " missing";                                           \
        return false;                                                          \
    }
        HY_NVRTC_SYM(CreateProgram)
        HY_NVRTC_SYM(DestroyProgram)
        HY_NVRTC_SYM(CompileProgram)
        HY_NVRTC_SYM(GetProgramLogSize)
        HY_NVRTC_SYM(GetProgramLog)
        HY_NVRTC_SYM(GetCUBINSize)
        HY_NVRTC_SYM(GetCUBIN)
        HY_NVRTC_SYM(AddNameExpression)
        HY_NVRTC_SYM(GetLoweredName)
        HY_NVRTC_SYM(Version)
        HY_NVRTC_SYM(GetErrorString)
#undef HY_NVRTC_SYM
        loaded = true;
        return true;
    }
};
inline Nvrtc &nvrtc()
{
    static Nvrtc n;
    return n;
}

// directory of libhy_cuda.so (the kernel headers are read from there; the cache lives below it)
inline std::string lib_dir()
{
    Dl_info info{};
    if (dladdr(reinterpret_cast<const void *>(&lib_dir), &info) && info.dli_fname) {
        std::string p(info.dli_fname);
        const size_t s = p.rfind('/');
        return s == std::string::npos ? std::string(".") : p.substr(0, s);
    }
    return ".";
}
inline std::string cache_dir()
{
    const char *e = std::getenv("HY_CUDA_JIT_CACHE");
    return (e && *e) ? std::string(e) : lib_dir() + "/jit_cache";
}

inline uint64_t fnv1a(const std::string &s, uint64_t h = 1469598103934665603ULL)
{
    for (unsigned char c : s) {
        h ^= c;
        h *= 1099511628211ULL;
    }
    return h;
}

inline std::string read_file(const std::string &path)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) return "";
    std::ostringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

// A compiled kernel: cubin image + lowered (mangled) name of the kernel in it.
struct Image {
    std::vector<char> cubin;
    std::string name;
    bool from_cache = false;
    double compile_s = 0;
};

// Stubs of the system headers the kernel sources include (NVRTC has none of them).
inline const char *stub_stdint()
{
    return "#pragma once\n"
           "typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;\n"
           "typedef int int32_t; typedef unsigned int uint32_t; typedef long int64_t; typedef unsigned long uint64_t;\n"
           "typedef unsigned long uintptr_t; typedef long intptr_t;\n";
}

// The headers of the kernel translation unit: their content is part of the cache key.
inline const std::vector<std::string> &kernel_headers()
{
    static const std::vector<std::string> v = {"hy_kernels.cuh",    "hy_devprog.h",     "hy_events.cuh", "hy_evtape.cuh",
                                               "hy_nbody_reg.cuh",  "hy_cr3bp_reg.cuh", "../../include/hy_cuda.h"};
    return v;
}

inline std::string options_string(const std::vector<std::string> &opts)
{
    std::string s;
    for (auto &o : opts) s += o + " ";
    return s;
}

// Compile `src` (or fetch it from the cache).  Returns "" on success.
inline std::string build(const std::string &src, const std::string &kname, Image &out)
{
    const std::string dir = lib_dir();
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device", "-I" + dir,
                                     "-I" + dir + "/../../include"};
    // cache key: generated source + kernel headers + options + compiler version
    uint64_t h = fnv1a(src);
    for (auto &hn : kernel_headers()) {
        const std::string body = read_file(dir + "/" + hn);
        if (body.empty()) return "cannot read kernel header " + dir + "/" + hn;
        h = fnv1a(body, h);
    }
    h = fnv1a(options_string(opts) + kname, h);
    Nvrtc &nv = nvrtc();
    int vmaj = 0, vmin = 0;
    const bool have_nvrtc = nv.load();
    if (have_nvrtc) nv.Version(&vmaj, &vmin);
    char keybuf[64];
    std::snprintf(keybuf, sizeof keybuf, "%016" PRIx64, h);
    const std::string cdir = cache_dir();
    const std::string cpath = cdir + "/" + keybuf + ".hyjit";
    {
        const std::string blob = read_file(cpath);
        if (blob.size() > 12 && std::memcmp(blob.data(), "HYJ1", 4) == 0) {
            uint32_t nl = 0;
            std::memcpy(&nl, blob.data() + 4, 4);
            if (8 + (size_t)nl < blob.size()) {
                out.name.assign(blob.data() + 8, nl);
                out.cubin.assign(blob.begin() + 8 + nl, blob.end());
                out.from_cache = true;
                return "";
            }
        }
    }
    if (!have_nvrtc) return "NVRTC unavailable (" + nv.err + ") and no cached kernel " + cpath;
    static std::mutex mtx; // (NVRTC is thread-safe, but one compilation at a time keeps the memory bounded)
    std::lock_guard<std::mutex> lk(mtx);
    timespec t0{}, t1{};
    clock_gettime(CLOCK_MONOTONIC, &t0);
    void *prog = nullptr;
    const char *hdr_names[] = {"stdint.h", "cstdint", "stddef.h", "cuda_runtime.h"};
    const std::string stddef_stub = "#pragma once\ntypedef unsigned long size_t;\n";
    const std::string cstdint_stub = std::string("#pragma once\n#include <stdint.h>\n");
    const char *hdr_src[] = {stub_stdint(), cstdint_stub.c_str(), stddef_stub.c_str(), "#pragma once\n"};
    int rc = nv.CreateProgram(&prog, src.c_str(), "hy_jit_kernel.cu", 4, hdr_src, hdr_names);
    if (rc) return std::string("nvrtcCreateProgram: ") + nv.GetErrorString(rc);
    nv.AddNameExpression(prog, kname.c_str());
    std::vector<const char *> copts;
    for (auto &o : opts) copts.push_back(o.c_str());
    rc = nv.CompileProgram(prog, (int)copts.size(), copts.data());
    if (rc) {
        size_t ls = 0;
        nv.GetProgramLogSize(prog, &ls);
        std::string log(ls, '\0');
        if (ls) nv.GetProgramLog(prog, &log[0]);
        nv.DestroyProgram(&prog);
        if (const char *dump = std::getenv("HY_CUDA_JIT_DUMP")) {
            std::ofstream f(std::string(dump) + "/hy_jit_failed.cu");
            f << src;
        }
        if (log.size() > 4000) log.resize(4000);
        return std::string("NVRTC compilation failed: ") + nv.GetErrorString(rc) + "\n" + log;
    }
    const char *lowered = nullptr;
    rc = nv.GetLoweredName(prog, kname.c_str(), &lowered);
    if (rc || !lowered) {
        nv.DestroyProgram(&prog);
        return "nvrtcGetLoweredName failed";
    }
    out.name = lowered;
    size_t cs = 0;
    nv.GetCUBINSize(prog, &cs);
    out.cubin.resize(cs);
    nv.GetCUBIN(prog, out.cubin.data());
    nv.DestroyProgram(&prog);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    out.compile_s = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    out.from_cache = false;
    if (const char *dump = std::getenv("HY_CUDA_JIT_DUMP")) {
        std::ofstream f(std::string(dump) + "/hy_jit_" + keybuf + ".cu");
        f << src;
    }
    // store (best effort; atomic rename so that concurrent processes never see a partial file)
    ::mkdir(cdir.c_str(), 0777);
    {
        const std::string tmp = cpath + ".tmp" + std::to_string((long)::getpid());
        std::ofstream f(tmp, std::ios::binary);
        if (f) {
            const uint32_t nl = (uint32_t)out.name.size();
            f.write("HYJ1", 4);
            f.write(reinterpret_cast<const char *>(&nl), 4);
            f.write(out.name.data(), nl);
            f.write(out.cubin.data(), (std::streamsize)out.cubin.size());
            f.close();
            if (::rename(tmp.c_str(), cpath.c_str()) != 0) ::unlink(tmp.c_str());
        }
    }
    return "";
}

// --------------------------------------------------------------------------------------------
// Driver API entry points (through the runtime: no link-time dependency on libcuda)
// --------------------------------------------------------------------------------------------
struct Driver {
    int (*ModuleLoadData)(void **, const void *) = nullptr;
    int (*ModuleUnload)(void *) = nullptr;
    int (*ModuleGetFunction)(void **, void *, const char *) = nullptr;
    int (*FuncSetAttribute)(void *, int, int) = nullptr;
    int (*FuncGetAttribute)(int *, int, void *) = nullptr;
    int (*LaunchKernel)(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void *, void **, void **) = nullptr;
    std::string err;
    std::mutex mtx;
    bool ok = false;
    bool load()
    {
        std::lock_guard<std::mutex> lk(mtx);
        if (ok) return true;
        auto get = [&](const char *nm, void **fp) {
            cudaDriverEntryPointQueryResult qr;
            if (cudaGetDriverEntryPoint(nm, fp, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess || !*fp) {
                err = std::string("driver entry point ") + nm + " not found";
                return false;
            }
            return true;
        };
        ok = get("cuModuleLoadData", (void **)&ModuleLoadData) && get("cuModuleUnload", (void **)&ModuleUnload) &&
             get("cuModuleGetFunction", (void **)&ModuleGetFunction) && get("cuFuncSetAttribute", (void **)&FuncSetAttribute) &&
             get("cuFuncGetAttribute", (void **)&FuncGetAttribute) && get("cuLaunchKernel", (void **)&LaunchKernel);
        return ok;
    }
};
inline Driver &driver()
{
    static Driver d;
    return d;
}

// A kernel loaded into the current device's primary context.
struct Loaded {
    void *module = nullptr;
    void *func = nullptr;
    int regs = 0;
};

inline std::string load(const Image &img, uint32_t smem_bytes, bool prefer_l1, Loaded &out)
{
    Driver &dr = driver();
    if (!dr.load()) return dr.err;
    cudaFree(nullptr); // make sure the primary context exists and is current
    int rc = dr.ModuleLoadData(&out.module, img.cubin.data());
    if (rc) return "cuModuleLoadData failed (" + std::to_string(rc) + ")";
    rc = dr.ModuleGetFunction(&out.func, out.module, img.name.c_str());
    if (rc) return "cuModuleGetFunction(" + img.name + ") failed (" + std::to_string(rc) + ")";
    // CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES = 8, PREFERRED_SHARED_MEMORY_CARVEOUT = 9, NUM_REGS = 4
    rc = dr.FuncSetAttribute(out.func, 8, (int)smem_bytes);
    if (rc) return "cuFuncSetAttribute(max dynamic shared memory) failed (" + std::to_string(rc) + ")";
    if (prefer_l1) dr.FuncSetAttribute(out.func, 9, 0); // jets in global memory: all of L1 for them
    dr.FuncGetAttribute(&out.regs, 4, out.func);
    return "";
}

inline void unload(Loaded &l)
{
    if (l.module && driver().ok) driver().ModuleUnload(l.module);
    l.module = l.func = nullptr;
}

template <typename KP>
inline cudaError_t launch(const Loaded &l, const KP &P, uint32_t ctas, uint32_t threads, uint32_t smem, cudaStream_t s)
{
    void *args[] = {const_cast<KP *>(&P)};
    const int rc = driver().LaunchKernel(l.func, ctas, 1, 1, threads, 1, 1, smem, (void *)s, args, nullptr);
    return rc == 0 ? cudaSuccess : cudaErrorLaunchFailure;
}

} // namespace jit
} // namespace hy
