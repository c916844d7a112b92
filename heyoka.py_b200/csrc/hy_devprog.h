// hy_devprog.h - plain-data layout of the device-side program (ops, terms, reference flags).
// Shared by the host scheduler (hy_schedule.hpp), the kernels (hy_kernels.cuh) and the run-time
// compiled kernels (hy_jit.hpp: this header must stay free of host-only includes).
#pragma once
#include <stdint.h>

namespace hy {

// Device op: 16 bytes, one LDS.128.
struct DOp {
    uint8_t opcode; // hy_opcode, or OP_NOP
    uint8_t flags;  // bits 0-3: HY_OPF_*; bits 4-7: jet flag of dst, dst2, a, b
    uint16_t n;     // number of terms
    uint16_t dst, dst2;
    uint16_t a, b; // operand offsets; for term ops b = first term position in the lane's stream
    uint16_t imm;  // index into the immediate table
    uint16_t pad;
};
static_assert(sizeof(DOp) == 16, "DOp must be 16 bytes");
// Device term: 16 bytes.
struct DTerm {
    double coef;
    uint32_t src; // offset | bit 31 = jet
    uint32_t aux; // LINCOMB: offset of the multiplier (parameter row or 1.0); MULSH: dst ref
};
static_assert(sizeof(DTerm) == 16, "DTerm must be 16 bytes");

enum : uint8_t { OP_NOP = 255, DOP_PAIR = 64 };
// DOP_PAIR: fused "pair interaction" cluster (superinstruction)
//   d_i = (+-)A_i (+-)B_i  (i < n <= 3),  r2 = sum_i d_i^2,  w = r2^alpha,  t_i = d_i * w
// terms (2 per component, in the lane's stream at o.b):
//   [2i]   src = A_i ref, aux = B_i ref, coef = sign code (0:+a+b 1:-a+b 2:+a-b 3:-a-b)
//   [2i+1] src = d_i jet ref,  aux = t_i output ref
// o.a = r2 jet, o.dst = w jet, o.dst2 = scratch row holding 1/r2[0], o.imm = alpha.
enum : uint8_t { DF_JDST = 0x10, DF_JDST2 = 0x20, DF_JA = 0x40, DF_JB = 0x80 };
// Ping-pong references (DOp::pad bits / DTerm bit 30): a spilled state variable
// keeps only orders k and k+1 on chip, at base + (order & 1).
enum : uint16_t { DP_DST = 0x1, DP_A = 0x4, DP_B = 0x8, DP_NOPAR = 0x10 };
// LINCOMB terms: aux = multiplier offset (low 16 bits) | order mask (high 16 bits);
// the operand of the term at order k sits at (src & 0xffffff) + (k & mask):
// mask = 0xffff for jets, 1 for spilled (ping-pong) state variables, 0 for single rows.
#define HY_DREF_JET 0x80000000u
#define HY_DREF_PP 0x40000000u

} // namespace hy
