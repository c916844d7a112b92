// hy_cuda.cu - C ABI of libhy_cuda (see include/hy_cuda.h) over the sm_100a
// kernels in hy_kernels.cuh.  No CPU fallback exists: every entry point needs
// a CUDA device and fails loudly otherwise.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <string>
#include <vector>

#include "hy_kernels.cuh"
#include "hy_nbody_match.hpp"
#include "hy_cr3bp_match.hpp"
#include "hy_nb_launch.hpp"
#include "hy_evtape_host.hpp"
#include "hy_jit.hpp"

namespace {

thread_local std::string g_err;

int fail(const std::string &msg)
{
    g_err = msg;
    return 1;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" +     \
                        std::to_string(__LINE__) + ")");                                                 \
    } while (0)

} // namespace

// Continuous-output record: owns the chunk pool written by the propagate kernel
// (hy_kernels.cuh, RecDev) and everything needed to evaluate it.  Created by
// hy_propagate(c_output) inside a context, handed to the caller by hy_cout_detach.
struct hy_cout {
    int device = 0;
    size_t rb = 8;
    uint32_t B = 0, n = 0, P1 = 0;
    std::vector<void *> segs; // pool segments (device)
    void **d_seg = nullptr;   // device copy of the segment table (HY_REC_MAXSEG entries)
    uint32_t seg_chunks = 0, rec_len = 0, chunk_len = 0, si = 0, sk = 1;
    void *d_lane = nullptr; // one allocation: next | dir_next | head | tail | count | nchunks | dir_off | t0_hi | t0_lo
    unsigned int *d_next = nullptr, *d_dir_next = nullptr;
    uint32_t *d_head = nullptr, *d_tail = nullptr, *d_count = nullptr, *d_nch = nullptr, *d_dir_off = nullptr;
    void *d_t0hi = nullptr, *d_t0lo = nullptr;
    uint32_t *d_dir = nullptr;
    size_t dir_cap = 0;
    bool indexed = false;
    cudaStream_t stream = nullptr;
    void *d_tmp_in = nullptr, *d_tmp_out = nullptr;
    size_t tmp_in_bytes = 0, tmp_out_bytes = 0;
};
constexpr int HY_REC_MAXSEG = 64;

struct hy_ctx {
    int device = 0;
    int fp_bits = 64;
    size_t rb = 8; // bytes per real
    hy_dims d{};
    uint32_t B = 0;
    double tol = 0;
    int high_accuracy = 0;
    int no_jit = 0; // compact_mode: stay on the tape interpreter
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // tape (host copies: kept for re-scheduling and for hy_clone)
    std::vector<hy_op> h_ops;
    std::vector<hy_term> h_terms;
    std::vector<uint32_t> h_levels;
    hy::Program prog;
    void *d_prog = nullptr; // [ops | terms | imm]
    uint32_t *d_phase = nullptr;
    uint32_t *d_ev = nullptr;       // remapped event jet rows (device layout)
    std::vector<uint32_t> h_ev_ref; // ABI event references
    // events on the register-resident kernels: the ODE-only tape (for the matchers) and the event tape
    hy_dims d_ode{};
    std::vector<hy_op> h_ops_ode;
    std::vector<hy_term> h_terms_ode;
    std::vector<hy_op> h_evt_ops;
    std::vector<hy_term> h_evt_terms;
    std::vector<uint32_t> h_evt_ref, h_evt_start;
    uint32_t evt_rows = 0;
    bool have_evt = false; // the caller supplied both tapes
    bool use_evt = false;  // ... and the ODE tape was matched: the event tape is in use
    void *d_evt = nullptr;
    hy::EvtDev evt_dev{};
    std::vector<unsigned char> evt_blob;
    hy::EvtProgram evt_prog; // (kept for the event code generator, hy_jit.hpp)
    unsigned long long *d_evstats = nullptr;
    std::vector<int32_t> h_ev_dir;
    std::vector<double> h_ev_cd;
    uint32_t *d_srow = nullptr;
    uint32_t *d_recoff = nullptr; // register-resident kernels: record element -> column offset
    int32_t *d_ssp = nullptr;
    void *d_gjet = nullptr;
    // lanes (device)
    void *d_state = nullptr, *d_pars = nullptr, *d_thi = nullptr, *d_tlo = nullptr, *d_lasth = nullptr;
    void *d_tf = nullptr, *d_tfhi = nullptr, *d_tflo = nullptr, *d_mdt = nullptr, *d_minh = nullptr, *d_maxh = nullptr,
         *d_tc = nullptr;
    long long *d_outcome = nullptr;
    unsigned long long *d_nsteps = nullptr;
    unsigned int *d_counter = nullptr;
    unsigned char *d_active = nullptr;
    uint32_t *d_gidx = nullptr;
    void *d_gws = nullptr;
    void *d_evws = nullptr; // event workspace slab of the register-resident kernels (EvtDev::gws)
    // events
    int32_t *d_ev_dir = nullptr;
    double *d_ev_cd = nullptr;
    void *d_cd_elapsed = nullptr, *d_cd_total = nullptr;
    hy_event_rec *d_log = nullptr;
    unsigned long long *d_log_count = nullptr;
    unsigned long long log_cap = 0;
    // angle reduction (device-side post-step op)
    uint32_t *d_red = nullptr;
    uint32_t n_red = 0;
    // continuous output: the record being written / a recycled one
    hy_cout *rec = nullptr, *rec_spare = nullptr;
    // scratch for grid / dense evaluation
    void *d_tmp_in = nullptr, *d_tmp_out = nullptr;
    size_t tmp_in_bytes = 0, tmp_out_bytes = 0;
    // launch geometry
    hy_launch_info li{};
    uint32_t TS = 0;
    // run-time compiled kernel (hy_jit.hpp; li.kernel_variant == HY_VARIANT_JIT)
    hy::jit::Image jit_img;
    hy::jit::Loaded jit_k;
    // N-body tapes with parametric masses: the register-resident kernel built at hy_create time with
    // HY_NBR_PAR - the plain build (jit_img / jit_k hold the FX build)
    hy::jit::Image jit_img_plain;
    hy::jit::Loaded jit_k_plain;
    std::string jit_defs; // macro definitions every run-time build of this context's kernel starts with
    // timing
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_ms = 0;
    uint64_t last_launches = 0;
};

namespace {

template <typename R, int G, bool SMEM, int NB = 0>
cudaError_t launch_g(const hy::KParams<R> &P, const hy_launch_info &li, cudaStream_t s)
{
    auto kern = hy::propagate_kernel<R, G, SMEM, NB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)li.smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<li.ctas, li.threads, li.smem_bytes, s>>>(P);
    return cudaGetLastError();
}

template <typename R>
cudaError_t launch(const hy::KParams<R> &P, const hy_launch_info &li, cudaStream_t s, bool fx, const hy::jit::Loaded *jk,
                   const hy::jit::Loaded *jkp = nullptr)
{
    if (!fx && jkp && jkp->func && li.kernel_variant != HY_VARIANT_JIT)
        return hy::jit::launch(*jkp, P, li.ctas, li.threads, li.smem_bytes, s);
    // register-resident N-body kernels (hy_nbody_reg.cuh)
    // a register-resident kernel rebuilt at hy_create time with its event functions as generated code
    if (fx && jk && jk->func && li.kernel_variant != HY_VARIANT_JIT)
        return hy::jit::launch(*jk, P, li.ctas, li.threads, li.smem_bytes, s);
    switch (li.kernel_variant) { // instantiated in hy_nb3.cu ... hy_nb6.cu
    case 0: break;
    case HY_VARIANT_JIT: // generated from the tape and compiled at hy_create time (hy_jit.hpp)
        if (!jk || !jk->func) return cudaErrorInvalidValue;
        return hy::jit::launch(*jk, P, li.ctas, li.threads, li.smem_bytes, s);
    case 3: return hy::launch_nbody_kernel<R, 3>(P, li, s, fx);
    case 4: return hy::launch_nbody_kernel<R, 4>(P, li, s, fx);
    case 5: return hy::launch_nbody_kernel<R, 5>(P, li, s, fx);
    case 6: return hy::launch_nbody_kernel<R, 6>(P, li, s, fx);
    case hy::NBR_VARIANT_P22: // 6 bodies, unrolled to order 22 (hy_nb6.cu)
        if constexpr (std::is_same<R, double>::value) return hy::launch_nbody_kernel_p22(P, li, s, fx);
        return cudaErrorInvalidValue;
    case 106: // warpgroup rotation (experimental: the plain build only)
        if constexpr (std::is_same<R, double>::value) return fx ? cudaErrorNotSupported : hy::launch_nbody_kernel_wgx(P, li, s);
        return cudaErrorInvalidValue;
    case hy::CRB_VARIANT: return hy::launch_cr3bp_kernel<R>(P, li, s, fx); // hy_cr3bp.cu
    case hy::CRB_VARIANT_P22:
        if constexpr (std::is_same<R, double>::value) return hy::launch_cr3bp_kernel_p22(P, li, s, fx);
        return cudaErrorInvalidValue;
    default: return cudaErrorInvalidValue;
    }
    if (li.ws_in_smem) {
        switch (li.group) {
        case 1: return launch_g<R, 1, true>(P, li, s);
        case 4: return launch_g<R, 4, true>(P, li, s);
        case 16: return launch_g<R, 16, true>(P, li, s);
        default: return cudaErrorInvalidValue;
        }
    }
    // Global-memory workspace fallback (jets too large for shared memory).
    switch (li.group) {
    case 1: return launch_g<R, 1, false>(P, li, s);
    default: return cudaErrorInvalidValue;
    }
}

template <typename R, int G, bool SMEM, int NB = 0> int regs_of()
{
    cudaFuncAttributes a{};
    if (cudaFuncGetAttributes(&a, hy::propagate_kernel<R, G, SMEM, NB>) != cudaSuccess) return 0;
    return a.numRegs;
}

template <typename R> int regs_for_group(uint32_t g, bool smem, uint32_t variant)
{
    switch (variant) {
    case 106: return 168;
    case hy::CRB_VARIANT: return hy::regs_cr3bp_kernel<R>();
    case hy::CRB_VARIANT_P22: return hy::regs_cr3bp_kernel_p22();
    case 3: return hy::regs_nbody_kernel<R, 3>();
    case 4: return hy::regs_nbody_kernel<R, 4>();
    case 5: return hy::regs_nbody_kernel<R, 5>();
    case 6: return hy::regs_nbody_kernel<R, 6>();
    case hy::NBR_VARIANT_P22: return hy::regs_nbody_kernel_p22();
    default: break;
    }
    if (!smem) return regs_of<R, 1, false>();
    switch (g) {
    case 1: return regs_of<R, 1, true>();
    case 4: return regs_of<R, 4, true>();
    default: return regs_of<R, 16, true>();
    }
}

uint32_t env_u32(const char *name, uint32_t dflt)
{
    const char *v = std::getenv(name);
    if (!v || !*v) return dflt;
    return (uint32_t)std::strtoul(v, nullptr, 10);
}

hy::ProgDims prog_dims(const hy::Program &p)
{
    return hy::ProgDims{p.n_slots, p.n_tslots, (uint32_t)p.imm.size(), p.n_phases, p.ws_len,
                        p.par_off, p.one_off,  p.n_spill,              p.evt_bytes};
}

// Plan, generate and compile (or fetch from the cache) the run-time compiled kernel of a tape.
// Returns 1: `pr` / `img` hold the kernel (row offsets of `pr` pre-multiplied for the interleaved
// workspace), 0: the tape is served better by the interpreter, -1: failure (`err`).
int jit_plan(const hy_dims &d, const hy_op *ops, const hy_term *terms, const uint32_t *ev_ref, int fp_bits, uint32_t B,
             uint32_t n_sm, uint32_t smem_optin, bool force, hy::Program &pr, bool &smem, uint32_t &T, hy::jit::Image &img,
             std::string &err)
{
    const uint32_t rb = (uint32_t)fp_bits / 8u;
    err = hy::build_program(d, ops, terms, ev_ref, 1, false, pr, false);
    if (!err.empty()) return -1;
    // order blocking of the products (HY_CUDA_JIT_BLOCK: block size; default off - the generated code
    // then rounds exactly like the interpreter.  Blocking halves the DRAM traffic of config 4 but the
    // sweep is bound by dependent memory round trips, not by bandwidth: 8.4e6 steps/s vs 1.17e7)
    hy::jit::Gen gen(d, pr, env_u32("HY_CUDA_JIT_BLOCK", 0));
    pr.ws_len += gen.q_rows;
    gen.pf_dist = env_u32("HY_CUDA_JIT_PF_DIST", 0);
    gen.pf_level = env_u32("HY_CUDA_JIT_PF_LEVEL", 1);
    gen.batch = env_u32("HY_CUDA_JIT_BATCH", 64);
    // Products / quotients of a level that share an operand load it once (convn_wide, jop_divsh; bit-identical).
    // Measured on config 4 (10^6 lanes, 512-thread CTAs = 128 registers): 30 % fewer operand loads but 12-term blocks
    // instead of whole convolutions in flight - 1.09e7 against 1.51e7 steps/s.  Off by default.
    gen.div_group = env_u32("HY_CUDA_JIT_SHARE", 0) != 0;
    const size_t col_bytes = (size_t)pr.ws_len * rb;
    if (!force && col_bytes <= env_u32("HY_CUDA_JIT_MIN_BYTES", 3072)) return 0;
    hy::ProgDims pd0 = prog_dims(pr);
    pd0.n_slots = pd0.n_tslots = 0; // (the program is code: no ops / terms are staged)
    hy::SmemLayout L0 = hy::make_layout(d, pd0, 1, 0, pr.ws_len, rb, 0);
    const uint32_t fixed = L0.total + 64;
    const uint32_t fit = fixed < smem_optin ? (uint32_t)((smem_optin - fixed) / col_bytes) & ~31u : 0u;
    smem = fit >= env_u32("HY_CUDA_JIT_SMEM_MIN_THREADS", 128);
    T = smem ? std::min(fit, 512u) : (std::max(32u, env_u32("HY_CUDA_JIT_THREADS", 512)) & ~31u);
    if (!smem && !env_u32("HY_CUDA_JIT_THREADS", 0)) {
        // A trajectory is one long indivisible job of a thread: with B trajectories on n_sm * T thread slots
        // the last "wave" is partly empty (125 000 trajectories on 148 x 512 slots: 1.65 waves, 82 % of the
        // slots busy).  Throughput grows like sqrt(T) between 256 and 512 threads (measured on config 4):
        // pick the CTA size that maximises sqrt(T) x slot efficiency.
        double best = -1;
        for (uint32_t t = 256; t <= 512; t += 32) {
            const double slots = (double)n_sm * t, waves = std::ceil((double)B / slots);
            const double score = std::sqrt((double)t) * (double)B / (waves * slots);
            if (score > best * 1.0001) {
                best = score;
                T = t;
            }
        }
    }
    // do not keep more trajectories resident than the batch can feed
    const uint32_t need = std::max(1u, (B + n_sm - 1) / n_sm);
    T = std::max(32u, std::min(T, (need + 31u) & ~31u));
    const std::string src = gen.source(fp_bits, smem, T);
    if (src.empty()) {
        err = "op without a generator";
        return -1;
    }
    err = hy::jit::build(src, hy::jit::kernel_name(fp_bits, smem), img);
    if (!err.empty()) return -1;
    // the kernel indexes the interleaved workspace in elements: pre-multiply the row offsets
    for (auto &r : pr.state_row) r *= hy::jit::WS;
    for (auto &r : pr.ev_ref) r *= hy::jit::WS;
    pr.par_off *= hy::jit::WS;
    pr.one_off *= hy::jit::WS;
    pr.n_slots = pr.n_tslots = 0; // nothing to stage: the program is code
    pr.ops.clear();
    pr.terms.clear();
    pr.n_phases = 0;
    pr.phase_slot = {0};
    return 1;
}

// Name of the FX build of a register-resident kernel that carries generated event functions ("" for a
// variant that has none).
static std::string evt_kernel_name(uint32_t variant, int fp_bits)
{
    const char *R = fp_bits == 64 ? "double" : "float";
    if (variant == (uint32_t)hy::CRB_VARIANT || variant == (uint32_t)hy::CRB_VARIANT_P22)
        return std::string("hy::propagate_kernel<") + R + ", 2, true, -1, false, " +
               (variant == (uint32_t)hy::CRB_VARIANT_P22 ? "hy::CRB_PMAX_HI" : "hy::NBR_PMAX") + ", true>";
    if (variant >= 3 && variant <= 8)
        return std::string("hy::propagate_kernel<") + R + (variant > 6 ? ", 32, true, " : ", 16, true, ") + std::to_string(variant) +
               ", false, hy::NBR_PMAX, true>";
    if (variant == (uint32_t)hy::NBR_VARIANT_P22) return "hy::propagate_kernel<double, 16, true, 6, false, hy::NBR_LMAX, true>";
    return "";
}

int upload_program(hy_ctx *c);

// Choose the launch geometry for a tape: group size G (threads cooperating on
// one trajectory), trajectories per CTA T; schedule the tape for that G and
// upload the program.
int choose_geometry(hy_ctx *c)
{
    cudaDeviceProp prop{};
    CU(cudaGetDeviceProperties(&prop, c->device));
    int smem_optin = 0;
    CU(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    const hy_dims &d = c->d;
    const uint32_t max_threads = 256; // == __launch_bounds__ of the 255-register kernels (hy_max_threads)
    hy_launch_info &li = c->li;
    li.n_sm = (uint32_t)prop.multiProcessorCount;
    const bool force_global = env_u32("HY_CUDA_FORCE_GLOBAL_WS", 0) != 0;
    const uint32_t Genv = env_u32("HY_CUDA_GROUP", 0);

    double best_score = -1;
    uint32_t bestG = 0, bestT = 0, bestRS = 0;
    bool best_smem = false;
    hy::Program best;
    li.kernel_variant = 0;
    // Register-resident kernel for N-body tapes (hy_nbody_reg.cuh): the jets of the pair
    // interactions live in registers, the state jets + a small exchange buffer in shared memory.
    // With events the matchers look at the ODE-only tape; the event functions then run from the
    // event tape (hy_evtape.cuh).  Without the two extra tapes an event-carrying system stays on the
    // interpreter (the matchers reject n_events != 0).
    const bool evt_ok = c->have_evt && env_u32("HY_CUDA_NO_REG_EVENTS", 0) == 0;
    const hy_dims &md = evt_ok ? c->d_ode : d;
    const hy_op *mops = evt_ok ? c->h_ops_ode.data() : c->h_ops.data();
    const hy_term *mterms = evt_ok ? c->h_terms_ode.data() : c->h_terms.data();
    c->use_evt = false;
    std::string evt_err;
    // Append the event workspace / interval scratch to the column of a matched tape.
    // (in_global: the workspace goes to a slab in global memory instead - the column keeps the state jets
    //  only, more trajectories fit in shared memory; see EvtDev::gws)
    auto attach_events = [&](hy::Program &pr, bool in_global) -> bool {
        if (!evt_ok) return true;
        hy::EvtProgram ep;
        evt_err = hy::build_event_program(d.n_state, d.order, c->h_evt_ops, c->h_evt_terms, c->h_evt_ref,
                                          c->h_evt_start, c->evt_rows, ep);
        if (!evt_err.empty()) return false;
        const uint32_t ews_off = in_global ? 0u : (pr.ws_len + 1u) & ~1u;
        const uint32_t eiv_off = ews_off + ((c->evt_rows + 1u) & ~1u);
        const uint32_t end = eiv_off + 2u * (d.n_state + ep.n_slots);
        if (!in_global) {
            pr.ws_len = end;
            pr.par_off = pr.one_off = pr.ws_len;
        }
        pr.ev_ref.clear();
        for (uint32_t off : ep.ev_off) pr.ev_ref.push_back(ews_off + off);
        pr.evt_bytes = (uint32_t)ep.blob.size();
        c->evt_blob = ep.blob;
        c->evt_prog = ep;
        c->evt_dev = hy::EvtDev{nullptr, ep.n_ops, ep.n_terms, ep.n_imm, d.n_events, ews_off, eiv_off, ep.n_slots,
                                (uint32_t)ep.blob.size(), nullptr, nullptr, in_global ? (end + 3u) & ~3u : 0u};
        return true;
    };
    // HY_CUDA_EVT_GLOBAL_WS = 0: always inside the column, 1: always the global slab, 2 (default): the slab when
    // it lets more trajectories be resident
    const uint32_t evt_gws_mode = env_u32("HY_CUDA_EVT_GLOBAL_WS", 2);
    hy::NbMatch nbm;
    c->jit_defs.clear();
    c->jit_img_plain = hy::jit::Image();
    auto nbody_kname = [&](uint32_t variant, bool fxb) {
        const char *R = c->fp_bits == 64 ? "double" : "float";
        if (variant == (uint32_t)hy::NBR_VARIANT_P22)
            return std::string("hy::propagate_kernel<double, 16, true, 6, false, hy::NBR_LMAX, ") + (fxb ? "true>" : "false>");
        return std::string("hy::propagate_kernel<") + R + (variant > 6 ? ", 32, true, " : ", 16, true, ") + std::to_string(variant) +
               ", false, hy::NBR_PMAX, " + (fxb ? "true>" : "false>");
    };
    bool nb_ok = !force_global && !Genv && env_u32("HY_CUDA_NO_NBODY_REG", 0) == 0 &&
                 hy::match_nbody(md, mops, mterms, nbm) && hy::nbody_kernel_variant(nbm.nb, d.order, c->fp_bits);
    if (nb_ok && (nbm.has_par || nbm.g32)) {
        // Masses scaled by runtime parameters (no precompiled kernel reads parameters), or 7 / 8 bodies (21 / 28
        // pairs: 32-lane groups, one trajectory per warp): build the matched kernel now (NVRTC, ~6 s per build,
        // cached) with HY_NBR_PAR / HY_NBR_G32.  No compiler: the tape goes the general way.
        nb_ok = !c->no_jit && env_u32("HY_CUDA_JIT", 2) != 0 && env_u32("HY_CUDA_WGX", 0) == 0;
        if (nb_ok) {
            const uint32_t v = hy::nbody_kernel_variant(nbm.nb, d.order, c->fp_bits);
            const std::string defs = std::string(nbm.has_par ? "#define HY_NBR_PAR 1\n" : "") +
                                     (nbm.g32 ? "#define HY_NBR_G32 1\n" : "");
            const std::string src = defs + "#include \"hy_kernels.cuh\"\n";
            hy::jit::Image fxi;
            std::string e1 = hy::jit::build(src, nbody_kname(v, false), c->jit_img_plain);
            if (e1.empty()) e1 = hy::jit::build(src, nbody_kname(v, true), fxi);
            if (e1.empty()) {
                c->jit_img = fxi;
                c->jit_defs = defs;
            } else {
                nb_ok = false;
                c->jit_img_plain = hy::jit::Image();
                if (env_u32("HY_CUDA_JIT_VERBOSE", 0))
                    std::fprintf(stderr, "hy_cuda: N-body kernel with parametric masses not built: %s\n", e1.c_str());
            }
        }
    }
    if (nb_ok) {
        hy::Program pr;
        const uint32_t NG = nbm.g32 ? 32u : 16u; // lanes of a group
        pr.G = NG;
        pr.n_phases = 0;
        pr.phase_slot = {0};
        pr.imm = nbm.imm;
        // experimental: warpgroup rotation (24 trajectories per SM, registers traded between warpgroups)
        const bool wgx = env_u32("HY_CUDA_WGX", 0) != 0 && nbm.nb == 6 && c->fp_bits == 64 &&
                         d.order == (uint32_t)hy::NBR_PMAX && c->B >= 24u * li.n_sm;
        pr.ws_len = nbm.g32 ? (uint32_t)hy::NbrHostLayout(true).ws : (uint32_t)hy::nbr_ws(wgx);
        pr.par_off = pr.one_off = pr.ws_len;
        pr.n_spill = 0;
        for (uint32_t i = 0; i < d.n_state; ++i) pr.state_row.push_back((uint32_t)hy::nbr_state_off((int)i));
        pr.state_spill.assign(d.n_state, -1);
        pr.n_clusters = nbm.n_pairs;
        pr.lane_utilisation = (double)nbm.n_pairs / (double)NG;
        const bool ev_att = attach_events(pr, evt_gws_mode == 1) && !(wgx && evt_ok);
        if (nbm.has_par) { // parameter rows of the trajectory, after everything else in the column
            pr.par_off = pr.ws_len;
            pr.ws_len += d.n_par;
            pr.one_off = pr.ws_len;
        }
        // column stride: even (16-byte aligned vectors); + 2 spreads the two trajectories of a warp over the banks
        const uint32_t RS = (pr.ws_len + 1u) / 2u * 2u + 2u;
        hy::SmemLayout L0 = hy::make_layout(d, prog_dims(pr), NG, 0, RS, (uint32_t)c->rb, 0);
        const uint32_t fixed = L0.total + 64;
        if (ev_att && fixed < (uint32_t)smem_optin && ((uint32_t)smem_optin - fixed) / (RS * (uint32_t)c->rb + 4u) >= 2) {
            c->use_evt = evt_ok;
            bestG = NG;
            bestT = std::min(((uint32_t)smem_optin - fixed) / (RS * (uint32_t)c->rb + 4u), (wgx ? 384u : max_threads) / NG) & ~1u;
            bestRS = RS;
            best_smem = true;
            best = pr;
            li.kernel_variant = hy::nbody_kernel_variant(nbm.nb, d.order, c->fp_bits);
            if (wgx) {
                if (bestT != 24u) return fail("hy_create: HY_CUDA_WGX needs 24 trajectories per CTA in shared memory");
                li.kernel_variant = 106;
            }
        }
    }
    if (!li.kernel_variant && !c->jit_img_plain.cubin.empty()) {
        // (built for parametric masses, but the column did not fit: the builds are not used)
        c->jit_img_plain = hy::jit::Image();
        c->jit_img = hy::jit::Image();
        c->jit_defs.clear();
    }
    // Register-resident CR3BP kernel (hy_cr3bp_reg.cuh): two lanes per trajectory, state jets
    // [order][variable] in shared memory.
    hy::CrbMatch crm;
    if (!li.kernel_variant && !force_global && !Genv && env_u32("HY_CUDA_NO_CR3BP_REG", 0) == 0 &&
        hy::match_cr3bp(md, mops, mterms, c->fp_bits, crm)) {
        hy::Program pr;
        pr.G = 2;
        pr.n_phases = 0;
        pr.phase_slot = {0};
        pr.imm = crm.imm;
        pr.ws_len = (uint32_t)hy::CRB_XS * (d.order + 1);
        pr.par_off = pr.one_off = pr.ws_len;
        pr.n_spill = 0;
        for (uint32_t i = 0; i < d.n_state; ++i) pr.state_row.push_back(i);
        pr.state_spill.assign(d.n_state, -1);
        pr.n_clusters = 2;
        pr.lane_utilisation = 1.0;
        // column stride = 2 * odd: the 16 lanes of a half-warp (8 trajectories x 2 lanes, three
        // elements apart) hit 16 different 64-bit banks
        const uint32_t mt = (uint32_t)hy::hy_max_threads(2, true, -1, (int)c->rb);
        uint32_t RS = 0, fixed = 0;
        // trajectories per CTA of a column layout (0: does not fit)
        auto crb_fit = [&](const hy::Program &q) -> uint32_t {
            RS = q.ws_len;
            while (RS % 4u != 2u) ++RS;
            hy::SmemLayout L0 = hy::make_layout(d, prog_dims(q), 2, 0, RS, (uint32_t)c->rb, 0);
            fixed = L0.total + 64;
            if (fixed >= (uint32_t)smem_optin) return 0u;
            return std::min(((uint32_t)smem_optin - fixed) / (RS * (uint32_t)c->rb + 4u), mt / 2u) & ~15u;
        };
        // event workspace: inside the column, or - when that keeps fewer trajectories resident - a global slab
        bool ev_att = true, ev_glob = evt_gws_mode == 1;
        if (evt_ok && evt_gws_mode == 2) {
            hy::Program a = pr, b = pr;
            ev_att = attach_events(a, false) && attach_events(b, true);
            if (ev_att) ev_glob = crb_fit(b) > crb_fit(a);
        }
        ev_att = ev_att && attach_events(pr, ev_glob);
        const uint32_t crbT = ev_att ? crb_fit(pr) : 0u;
        if (crbT >= 16) {
            c->use_evt = evt_ok;
            bestG = 2;
            bestT = crbT;
            bestRS = RS;
            best_smem = true;
            best = pr;
            li.kernel_variant = hy::cr3bp_kernel_variant(d.order, c->fp_bits);
        }
    }
    for (uint32_t G : {1u, 4u, 16u}) { // group sizes with compiled kernels
        if (li.kernel_variant) break;
        if (Genv && G != Genv) continue;
        hy::Program pr;
        std::string err = hy::build_program(d, c->h_ops.data(), c->h_terms.data(), c->h_ev_ref.data(), G, true, pr);
        if (!err.empty()) return fail("hy_create: " + err);
        // Per-trajectory workspace column, padded to an odd element count so that
        // lanes working on different trajectories hit different banks.
        const uint32_t RS = pr.ws_len | 1u;
        hy::SmemLayout L0 = hy::make_layout(d, prog_dims(pr), G, 0, RS, (uint32_t)c->rb, 0);
        const uint32_t fixed = L0.total + 64;
        if (fixed > (uint32_t)smem_optin) continue;
        const uint32_t budget = (uint32_t)smem_optin - fixed;
        uint32_t Tfit = budget / (RS * (uint32_t)c->rb + 4u);
        bool smem = Tfit >= 1 && !force_global;
        if (!smem && G != 1) continue; // the global-workspace fallback kernel exists for G = 1 only
        const uint32_t mt = (uint32_t)hy::hy_max_threads((int)G, smem, 0); // 512 for the small-group interpreter variants
        uint32_t T = smem ? std::min(Tfit, mt / G) : std::max(1u, 256u / G);
        const double threads = std::min<double>((double)T * G, mt);
        const double score = pr.lane_utilisation * threads * (smem ? 1.0 : 0.05);
        if (score > best_score * 1.02) {
            best_score = score;
            bestG = G;
            bestT = T;
            best_smem = smem;
            best = pr;
            bestRS = RS;
        }
    }
    // Run-time compiled kernel (hy_jit.hpp): one thread per trajectory, the order sweep generated from
    // the tape.  HY_CUDA_JIT = 0: never, 1: always, 2 (default): when the interpreter would run the
    // tape badly - its workspace does not fit in shared memory, or lane utilisation x resident threads
    // (the score above) is low: few trajectories per SM and levels too narrow / too mixed for the
    // group (config 4: 10 trajectories, 16-lane groups half idle).  Wide regular tapes (an N-body
    // system with parametric masses) stay on the interpreter, which is faster for them (3.5e7 vs
    // 2.5e7 steps/s).  hy_create's `compact_mode` flag (reference kwarg) selects the interpreter.
    bool jit_smem = false;
    uint32_t jit_T = 0;
    uint32_t jit_mode = c->no_jit ? 0u : env_u32("HY_CUDA_JIT", 2);
    // (the A/B switches that keep a matched tape "on the tape interpreter" mean exactly that)
    if (jit_mode == 2 && (env_u32("HY_CUDA_NO_NBODY_REG", 0) || env_u32("HY_CUDA_NO_CR3BP_REG", 0) ||
                          env_u32("HY_CUDA_NO_REG_EVENTS", 0)))
        jit_mode = 0;
    // (central-force tapes - N-body problems beyond the six bodies of the register kernels, fixed centres -
    //  run mostly on the fused, order-specialised pair op: the interpreter beats one thread per trajectory
    //  there whatever the score says - 7 / 8 bodies: 1.98e7 / 1.84e7 vs 1.66e7 / 1.25e7 steps/s)
    uint32_t n_pair_ops = 0;
    for (const hy::DOp &q : best.ops) n_pair_ops += q.opcode == hy::DOP_PAIR;
    const bool pair_dominated = bestG && best_smem && 4u * n_pair_ops >= best.n_clusters && n_pair_ops >= 3;
    const bool interp_weak = !bestG || !best_smem ||
                             (best_score < (double)env_u32("HY_CUDA_JIT_MAX_SCORE", 200) && !pair_dominated);
    if (!li.kernel_variant && !force_global && !Genv && (jit_mode == 1 || (jit_mode == 2 && interp_weak))) {
        hy::Program pr;
        std::string jerr;
        const int jr = jit_plan(d, c->h_ops.data(), c->h_terms.data(), c->h_ev_ref.data(), c->fp_bits, c->B, li.n_sm,
                                (uint32_t)smem_optin, jit_mode == 1, pr, jit_smem, jit_T, c->jit_img, jerr);
        if (jr > 0) {
            bestG = 1;
            bestT = jit_T;
            bestRS = pr.ws_len;
            best_smem = jit_smem;
            best = pr;
            li.kernel_variant = HY_VARIANT_JIT;
        } else if (jr < 0) {
            if (jit_mode == 1) return fail("hy_create: " + jerr);
            if (env_u32("HY_CUDA_JIT_VERBOSE", 0))
                std::fprintf(stderr, "hy_cuda: run-time compilation failed, using the tape interpreter: %s\n", jerr.c_str());
        }
    }
    if (!bestG) return fail("hy_create: the program does not fit in shared memory for any group size");
    uint32_t G = bestG, T = bestT;
    li.ws_in_smem = best_smem ? 1 : 0;
    uint32_t Tenv = env_u32("HY_CUDA_TRAJ_PER_CTA", 0);
    if (Tenv) T = std::min(Tenv, T);
    // Do not keep more trajectories resident than the batch can feed.
    uint32_t per_cta_needed = std::max(1u, (c->B + li.n_sm - 1) / li.n_sm);
    T = std::max(1u, std::min(T, per_cta_needed));
    if (li.kernel_variant == HY_VARIANT_JIT) T = jit_T; // whole warps, the CTA size the kernel was compiled for
    if (li.kernel_variant) T = (T + 1u) & ~1u; // whole warps: the two trajectories of a warp step in lockstep
    if (li.kernel_variant == 106) T = 24;      // whole warpgroups
    if (li.kernel_variant == (uint32_t)hy::CRB_VARIANT || li.kernel_variant == (uint32_t)hy::CRB_VARIANT_P22)
        T = (T + 15u) & ~15u; // whole warps (16 trajectories)
    li.group = G;
    li.traj_per_cta = T;
    li.threads = ((T * G + 31) / 32) * 32;
    uint32_t ctas = (c->B + T - 1) / T;
    li.ctas = std::max(1u, std::min(ctas, li.n_sm * env_u32("HY_CUDA_CTAS_PER_SM", 1)));
    if (li.kernel_variant == HY_VARIANT_JIT && !li.ws_in_smem)
        li.ctas = std::max(1u, std::min(ctas, li.n_sm * env_u32("HY_CUDA_JIT_CTAS_PER_SM", 1)));
    const uint32_t RS = bestRS;
    c->TS = RS;
    c->prog = best;
    hy::SmemLayout L = hy::make_layout(d, prog_dims(c->prog), G, T, RS, (uint32_t)c->rb, (int)li.ws_in_smem);
    li.smem_bytes = L.total;
    if (li.smem_bytes > (uint32_t)smem_optin) return fail("tape does not fit in shared memory");
    // Events on a register-resident kernel: rebuild the kernel (FX build) with the event functions as
    // generated code (hy_jit.hpp, EvtGen; NVRTC compiles the one instantiation in 5-7 s).
    // HY_CUDA_JIT_EVT = 0: never (the event tape is interpreted), 1 (default): for batches of at least
    // HY_CUDA_JIT_EVT_MIN_BATCH lanes, 2: always.  A failed compilation leaves the interpreted event tape
    // in place.
    if (c->use_evt && li.kernel_variant != HY_VARIANT_JIT && !c->no_jit) {
        // (the 5-7 s of compilation pay for themselves on large ensembles only: below
        //  HY_CUDA_JIT_EVT_MIN_BATCH lanes - default 4096 - the event tape is interpreted; HY_CUDA_JIT_EVT=2
        //  generates the code whatever the batch)
        uint32_t ej = env_u32("HY_CUDA_JIT_EVT", 1);
        if (ej == 1 && c->B < env_u32("HY_CUDA_JIT_EVT_MIN_BATCH", 4096)) ej = 0;
        const std::string kname = ej >= 1 ? evt_kernel_name(li.kernel_variant, c->fp_bits) : std::string();
        if (!kname.empty()) {
            const std::string src = hy::jit::evt_kernel_source(c->evt_prog, c->prog.state_row, d.order, c->jit_defs, G,
                                                               2u * (d.n_state + c->evt_prog.n_slots));
            hy::jit::Image img;
            std::string jerr = hy::jit::build(src, kname, img);
            if (jerr.empty()) {
                CU(cudaSetDevice(c->device));
                hy::jit::unload(c->jit_k);
                jerr = hy::jit::load(img, li.smem_bytes, false, c->jit_k);
                if (jerr.empty()) c->jit_img = img;
            }
            if (!jerr.empty() && env_u32("HY_CUDA_JIT_VERBOSE", 0))
                std::fprintf(stderr, "hy_cuda: event code generation failed, the event tape is interpreted: %s\n", jerr.c_str());
        }
    }
    hy::jit::unload(c->jit_k_plain);
    if (!c->jit_img_plain.cubin.empty() && li.kernel_variant != HY_VARIANT_JIT) {
        CU(cudaSetDevice(c->device));
        std::string lerr = hy::jit::load(c->jit_img_plain, li.smem_bytes, false, c->jit_k_plain);
        if (lerr.empty() && !c->jit_k.func) lerr = hy::jit::load(c->jit_img, li.smem_bytes, false, c->jit_k);
        if (!lerr.empty()) return fail("hy_create: " + lerr);
        li.regs_per_thread = (uint32_t)c->jit_k_plain.regs;
        return upload_program(c);
    }
    if (li.kernel_variant == HY_VARIANT_JIT) {
        CU(cudaSetDevice(c->device));
        hy::jit::unload(c->jit_k);
        const std::string lerr = hy::jit::load(c->jit_img, li.smem_bytes, !li.ws_in_smem, c->jit_k);
        if (!lerr.empty()) return fail("hy_create: " + lerr);
        li.regs_per_thread = (uint32_t)c->jit_k.regs;
        return upload_program(c);
    }
    li.regs_per_thread = (uint32_t)(c->fp_bits == 64 ? regs_for_group<double>(G, li.ws_in_smem, li.kernel_variant)
                                                     : regs_for_group<float>(G, li.ws_in_smem, li.kernel_variant));
    return upload_program(c);
}

// Upload the scheduled program of `c` (c->prog, c->li, c->TS) to its device.
int upload_program(hy_ctx *c)
{
    const hy_dims &d = c->d;
    hy_launch_info &li = c->li;
    const uint32_t G = li.group, T = li.traj_per_cta, RS = c->TS;
    hy::SmemLayout L = hy::make_layout(d, prog_dims(c->prog), G, T, RS, (uint32_t)c->rb, (int)li.ws_in_smem);
    // The program blob [ops | terms | imm] (same layout as in shared memory).
    {
        std::vector<unsigned char> blob(L.off_phase, 0);
        std::memcpy(blob.data() + L.off_ops, c->prog.ops.data(), c->prog.ops.size() * sizeof(hy::DOp));
        if (!c->prog.terms.empty())
            std::memcpy(blob.data() + L.off_terms, c->prog.terms.data(), c->prog.terms.size() * sizeof(hy::DTerm));
        std::memcpy(blob.data() + L.off_imm, c->prog.imm.data(), c->prog.imm.size() * 8);
        if (c->d_prog) cudaFree(c->d_prog);
        CU(cudaMalloc(&c->d_prog, std::max<size_t>(16, blob.size())));
        CU(cudaMemcpy(c->d_prog, blob.data(), blob.size(), cudaMemcpyHostToDevice));
        if (c->d_phase) cudaFree(c->d_phase);
        CU(cudaMalloc(&c->d_phase, c->prog.phase_slot.size() * 4));
        CU(cudaMemcpy(c->d_phase, c->prog.phase_slot.data(), c->prog.phase_slot.size() * 4, cudaMemcpyHostToDevice));
        for (void *p : {(void *)c->d_srow, (void *)c->d_ssp, (void *)c->d_ev, c->d_gjet})
            if (p) cudaFree(p);
        CU(cudaMalloc(&c->d_srow, d.n_state * 4));
        CU(cudaMemcpy(c->d_srow, c->prog.state_row.data(), d.n_state * 4, cudaMemcpyHostToDevice));
        if (c->d_recoff) cudaFree(c->d_recoff);
        c->d_recoff = nullptr;
        if (li.kernel_variant && li.kernel_variant != HY_VARIANT_JIT) {
            // order stride of the state jets in the column of the matched kernel
            const bool crb = li.kernel_variant == (uint32_t)hy::CRB_VARIANT || li.kernel_variant == (uint32_t)hy::CRB_VARIANT_P22;
            const uint32_t xs = crb ? (uint32_t)hy::CRB_XS : (uint32_t)hy::NBR_JS, P1 = d.order + 1;
            std::vector<uint32_t> tab((size_t)d.n_state * P1);
            for (uint32_t i = 0; i < d.n_state; ++i)
                for (uint32_t k = 0; k < P1; ++k) tab[(size_t)i * P1 + k] = c->prog.state_row[i] + k * xs;
            CU(cudaMalloc(&c->d_recoff, tab.size() * 4));
            CU(cudaMemcpy(c->d_recoff, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
        }
        CU(cudaMalloc(&c->d_ssp, d.n_state * 4));
        CU(cudaMemcpy(c->d_ssp, c->prog.state_spill.data(), d.n_state * 4, cudaMemcpyHostToDevice));
        CU(cudaMalloc(&c->d_ev, std::max<size_t>(1, d.n_events) * 4));
        if (d.n_events) CU(cudaMemcpy(c->d_ev, c->prog.ev_ref.data(), d.n_events * 4, cudaMemcpyHostToDevice));
        const size_t gj_bytes = (size_t)li.ctas * T * std::max<uint32_t>(1, c->prog.n_spill) * (d.order + 1) * c->rb;
        CU(cudaMalloc(&c->d_gjet, gj_bytes));
        CU(cudaMemset(c->d_gjet, 0, gj_bytes));
    }
    if (!li.ws_in_smem) {
        if (c->d_gws) cudaFree(c->d_gws);
        CU(cudaMalloc(&c->d_gws, (size_t)li.ctas * T * RS * c->rb));
    }
    if (c->d_evt) cudaFree(c->d_evt);
    c->d_evt = nullptr;
    if (c->use_evt) {
        CU(cudaMalloc(&c->d_evt, c->evt_blob.size()));
        CU(cudaMemcpy(c->d_evt, c->evt_blob.data(), c->evt_blob.size(), cudaMemcpyHostToDevice));
        c->evt_dev.blob = c->d_evt;
        c->evt_dev.stats = nullptr;
        if (c->d_evws) cudaFree(c->d_evws);
        c->d_evws = nullptr;
        c->evt_dev.gws = nullptr;
        if (c->evt_dev.gstride) {
            const size_t bytes = (size_t)li.ctas * T * c->evt_dev.gstride * c->rb;
            CU(cudaMalloc(&c->d_evws, bytes));
            CU(cudaMemset(c->d_evws, 0, bytes));
            c->evt_dev.gws = c->d_evws;
        }
        if (env_u32("HY_CUDA_EVENT_STATS", 0)) {
            if (!c->d_evstats) CU(cudaMalloc((void **)&c->d_evstats, 16));
            CU(cudaMemset(c->d_evstats, 0, 16));
            c->evt_dev.stats = c->d_evstats;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------
// Continuous-output record (hy_cout): chunk pool + per-lane lists.
// ---------------------------------------------------------------------------
void rec_free(hy_cout *r)
{
    if (!r) return;
    cudaSetDevice(r->device);
    for (void *p : r->segs)
        if (p) cudaFree(p);
    for (void *p : {(void *)r->d_seg, r->d_lane, (void *)r->d_dir, r->d_tmp_in, r->d_tmp_out})
        if (p) cudaFree(p);
    if (r->stream) cudaStreamDestroy(r->stream);
    delete r;
}

int rec_add_segment(hy_cout *r)
{
    if ((int)r->segs.size() >= HY_REC_MAXSEG) return fail("continuous output: too many pool segments");
    const size_t bytes = (size_t)r->seg_chunks * r->chunk_len * r->rb;
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    if (bytes + (64u << 20) > free_b)
        return fail("hy_propagate: continuous output needs another " + std::to_string(bytes >> 20) +
                    " MiB of device memory but only " + std::to_string(free_b >> 20) + " MiB are free");
    void *p = nullptr;
    CU(cudaMalloc(&p, bytes));
    r->segs.push_back(p);
    CU(cudaMemcpy(r->d_seg + (r->segs.size() - 1), &p, sizeof(void *), cudaMemcpyHostToDevice));
    return 0;
}

int rec_create(hy_ctx *c, hy_cout **out)
{
    hy_cout *r = new hy_cout();
    *out = r;
    r->device = c->device;
    r->rb = c->rb;
    r->B = c->B;
    r->n = c->d.n_state;
    r->P1 = c->d.order + 1;
    r->rec_len = r->n * r->P1 + 2u;
    r->chunk_len = 2u + hy::HY_REC_CH * r->rec_len;
    // record layout (RecDev::si, sk): order-major for the FP64 CR3BP kernel - its column in shared memory is the
    // record, copied out by TMA (HY_CUDA_REC_BULK=0: element copies into the tc layout, as on every other kernel)
    r->si = r->P1;
    r->sk = 1;
    const bool crb = c->li.kernel_variant == (uint32_t)hy::CRB_VARIANT || c->li.kernel_variant == (uint32_t)hy::CRB_VARIANT_P22;
    if (crb && c->rb == 8 && env_u32("HY_CUDA_REC_BULK", 1)) {
        r->si = 1;
        r->sk = r->n;
    }
    const size_t chunk_bytes = (size_t)r->chunk_len * r->rb;
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    // first segment: room for 4 chunks (32 steps) per lane, within 40 % of the free memory; at least 256 chunks
    size_t want = std::max<size_t>(256, (size_t)4 * r->B);
    want = std::min<size_t>(want, (size_t)(0.4 * (double)free_b / (double)chunk_bytes));
    r->seg_chunks = (uint32_t)std::max<size_t>(16, std::min<size_t>(want, 0x7fffffffu / HY_REC_MAXSEG));
    CU(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    CU(cudaMalloc((void **)&r->d_seg, HY_REC_MAXSEG * sizeof(void *)));
    CU(cudaMemset(r->d_seg, 0, HY_REC_MAXSEG * sizeof(void *)));
    const size_t B = r->B;
    const size_t lane_bytes = 16 + 5 * B * 4 + 2 * B * 8;
    CU(cudaMalloc(&r->d_lane, lane_bytes));
    char *q = (char *)r->d_lane;
    r->d_next = (unsigned int *)q;
    r->d_dir_next = (unsigned int *)(q + 8);
    q += 16;
    r->d_t0hi = q, q += B * 8;
    r->d_t0lo = q, q += B * 8;
    r->d_head = (uint32_t *)q, q += B * 4;
    r->d_tail = (uint32_t *)q, q += B * 4;
    r->d_count = (uint32_t *)q, q += B * 4;
    r->d_nch = (uint32_t *)q, q += B * 4;
    r->d_dir_off = (uint32_t *)q;
    return rec_add_segment(r);
}

// Empty the record (stream-ordered on `s`).
int rec_reset(hy_cout *r, cudaStream_t s)
{
    const size_t B = r->B;
    CU(cudaMemsetAsync(r->d_lane, 0, 16 + 2 * B * 8, s));                  // next, dir_next, t0
    CU(cudaMemsetAsync(r->d_head, 0xff, 2 * B * 4, s));                    // head, tail = NONE
    CU(cudaMemsetAsync(r->d_count, 0, 3 * B * 4, s));                      // count, nchunks, dir_off
    r->indexed = false;
    return 0;
}

template <typename R> hy::RecDev<R> rec_dev(const hy_cout *r, int on, int append)
{
    hy::RecDev<R> d{};
    if (!r) return d;
    d.seg = (R *const *)r->d_seg;
    d.seg_chunks = r->seg_chunks;
    d.cap_chunks = (uint32_t)(r->seg_chunks * r->segs.size());
    d.next = r->d_next;
    d.head = r->d_head;
    d.tail = r->d_tail;
    d.count = r->d_count;
    d.t0_hi = (R *)r->d_t0hi;
    d.t0_lo = (R *)r->d_t0lo;
    d.rec_len = r->rec_len;
    d.chunk_len = r->chunk_len;
    d.si = r->si;
    d.sk = r->sk;
    d.on = on;
    d.append = append;
    return d;
}

int rec_ensure_tmp(hy_cout *r, size_t in_bytes, size_t out_bytes)
{
    if (in_bytes > r->tmp_in_bytes) {
        if (r->d_tmp_in) cudaFree(r->d_tmp_in);
        r->d_tmp_in = nullptr;
        r->tmp_in_bytes = 0;
        CU(cudaMalloc(&r->d_tmp_in, in_bytes));
        r->tmp_in_bytes = in_bytes;
    }
    if (out_bytes > r->tmp_out_bytes) {
        if (r->d_tmp_out) cudaFree(r->d_tmp_out);
        r->d_tmp_out = nullptr;
        r->tmp_out_bytes = 0;
        CU(cudaMalloc(&r->d_tmp_out, out_bytes));
        r->tmp_out_bytes = out_bytes;
    }
    return 0;
}

// Build the chunk directory (once per recording).
int rec_index(hy_cout *r)
{
    if (r->indexed) return 0;
    CU(cudaSetDevice(r->device));
    unsigned int used = 0;
    CU(cudaMemcpyAsync(&used, r->d_next, 4, cudaMemcpyDeviceToHost, r->stream));
    CU(cudaStreamSynchronize(r->stream));
    const size_t cap = (size_t)r->seg_chunks * r->segs.size();
    const size_t need = std::max<size_t>(1, std::min<size_t>(used, cap));
    if (need > r->dir_cap) {
        if (r->d_dir) cudaFree(r->d_dir);
        r->d_dir = nullptr;
        r->dir_cap = 0;
        CU(cudaMalloc((void **)&r->d_dir, need * 4));
        r->dir_cap = need;
    }
    CU(cudaMemsetAsync(r->d_dir_next, 0, 4, r->stream));
    const unsigned th = 128, bl = (unsigned)((r->B + th - 1) / th);
    if (r->rb == 8)
        hy::rec_index_kernel<double><<<bl, th, 0, r->stream>>>(rec_dev<double>(r, 1, 0), r->d_dir_off, r->d_dir,
                                                               r->d_dir_next, r->B);
    else
        hy::rec_index_kernel<float><<<bl, th, 0, r->stream>>>(rec_dev<float>(r, 1, 0), r->d_dir_off, r->d_dir,
                                                              r->d_dir_next, r->B);
    CU(cudaGetLastError());
    r->indexed = true;
    return 0;
}

__global__ void mask_outcome_kernel(const long long *__restrict__ outcome, long long code,
                                    unsigned char *__restrict__ active, uint32_t B)
{
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < B) active[l] = outcome[l] == code ? 1 : 0;
}

struct RunArgs {
    int mode = hy::MODE_UNTIL;
    int backward = 0;
    uint64_t max_steps = 0;
    bool have_mdt = false;
    int write_tc = 0;
    int rec_on = 0, rec_append = 0;
    const void *grid = nullptr;
    void *gout = nullptr;
    uint32_t grid_k = 0;
    bool use_active = false;
    int resume = 0;
    int pause_on_nt = 0;
    uint64_t launch_steps = 0;
};

template <typename R> hy::KParams<R> make_params(hy_ctx *c, const RunArgs &a)
{
    hy::KParams<R> P{};
    P.d = c->d;
    P.prog = c->d_prog;
    P.phase_slot = c->d_phase;
    P.ev_ref = c->d_ev;
    P.state_row = c->d_srow;
    P.state_spill = c->d_ssp;
    P.gjet = (R *)c->d_gjet;
    P.pd = prog_dims(c->prog);
    P.state = (R *)c->d_state;
    P.pars = (const R *)c->d_pars;
    P.t_hi = (R *)c->d_thi;
    P.t_lo = (R *)c->d_tlo;
    P.last_h = (R *)c->d_lasth;
    P.tf_hi = (const R *)c->d_tfhi;
    P.tf_lo = (const R *)c->d_tflo;
    P.mdt = a.have_mdt ? (const R *)c->d_mdt : nullptr;
    P.active = a.use_active ? c->d_active : nullptr;
    P.gidx = c->d_gidx;
    P.resume = a.resume;
    P.pause_on_nt = a.pause_on_nt;
    P.launch_steps = a.launch_steps;
    P.rec_off = c->d_recoff;
    P.red_idx = c->d_red;
    P.n_red = c->n_red;
    P.rec = rec_dev<R>(c->rec, a.rec_on, a.rec_append);
    P.evt = c->use_evt ? c->evt_dev : hy::EvtDev{};
    P.outcome = c->d_outcome;
    P.min_h = (R *)c->d_minh;
    P.max_h = (R *)c->d_maxh;
    P.n_steps = c->d_nsteps;
    P.tc = (R *)c->d_tc;
    P.grid = (const R *)a.grid;
    P.gout = (R *)a.gout;
    P.grid_k = a.grid_k;
    P.ev.dir = c->d_ev_dir;
    P.ev.cooldown = c->d_ev_cd;
    P.ev.cd_elapsed = (R *)c->d_cd_elapsed;
    P.ev.cd_total = (R *)c->d_cd_total;
    P.ev.log = c->d_log;
    P.ev.log_count = c->d_log_count;
    P.ev.log_cap = c->log_cap;
    P.ev.tol = (R)c->tol;
    P.counter = c->d_counter;
    P.gws = (R *)c->d_gws;
    P.B = c->B;
    P.T = c->li.traj_per_cta;
    P.TS = c->TS;
    P.nb_tb_off = (uint32_t)hy::NBR_TB0;
    P.wgx_wgs = env_u32("HY_CUDA_WGX_WGS", 3);
    P.max_steps = a.max_steps;
    P.mode = a.mode;
    P.backward = a.backward;
    P.write_tc = a.write_tc;
    P.high_accuracy = c->high_accuracy;
    P.ws_in_smem = (int)c->li.ws_in_smem;
    const double p = (double)c->d.order;
    P.rhofac = (R)(std::exp(-7.0 / (10.0 * (p - 1.0))) / (M_E * M_E));
    P.inv_p = (R)(1.0 / p);
    P.inv_pm1 = (R)(1.0 / (p - 1.0));
    return P;
}

int ensure_tc(hy_ctx *c)
{
    if (!c->d_tc) {
        size_t bytes = (size_t)c->d.n_state * (c->d.order + 1) * c->B * c->rb;
        CU(cudaMalloc(&c->d_tc, bytes ? bytes : 8));
        CU(cudaMemsetAsync(c->d_tc, 0, bytes, c->stream));
    }
    return 0;
}

int ensure_tmp(hy_ctx *c, size_t in_bytes, size_t out_bytes)
{
    if (in_bytes > c->tmp_in_bytes) {
        if (c->d_tmp_in) cudaFree(c->d_tmp_in);
        c->d_tmp_in = nullptr;
        c->tmp_in_bytes = 0;
        CU(cudaMalloc(&c->d_tmp_in, in_bytes));
        c->tmp_in_bytes = in_bytes;
    }
    if (out_bytes > c->tmp_out_bytes) {
        if (c->d_tmp_out) cudaFree(c->d_tmp_out);
        c->d_tmp_out = nullptr;
        c->tmp_out_bytes = 0;
        CU(cudaMalloc(&c->d_tmp_out, out_bytes));
        c->tmp_out_bytes = out_bytes;
    }
    return 0;
}

// One launch of the persistent kernel, stream-ordered, NO host synchronisation
// (the caller synchronises once, after queueing the result copies).
int launch_once(hy_ctx *c, const RunArgs &a)
{
    CU(cudaMemsetAsync(c->d_counter, 0, sizeof(unsigned int), c->stream));
    // the plain build serves an uninterrupted propagate / step; anything else needs the FX build
    const bool fx = a.rec_on || a.use_active || a.resume || a.pause_on_nt || a.launch_steps || c->n_red || c->use_evt;
    cudaError_t e;
    if (c->fp_bits == 64)
        e = launch<double>(make_params<double>(c, a), c->li, c->stream, fx, &c->jit_k, &c->jit_k_plain);
    else
        e = launch<float>(make_params<float>(c, a), c->li, c->stream, fx, &c->jit_k, &c->jit_k_plain);
    if (e != cudaSuccess) return fail(std::string("kernel launch: ") + cudaGetErrorString(e));
    ++c->last_launches;
    return 0;
}

// Launch (and, while recording, re-launch the lanes that ran out of recorder
// pool after growing it).  Timing brackets all launches of the call.
int run_kernel(hy_ctx *c, RunArgs a)
{
    c->last_ms = 0;
    c->last_launches = 0;
    if (c->B == 0) return 0;
    if (a.write_tc && ensure_tc(c)) return 1;
    CU(cudaEventRecord(c->ev0, c->stream));
    if (launch_once(c, a)) return 1;
    while (a.rec_on) {
        // pool exhausted?  (4-byte read-back per recording launch)
        unsigned int used = 0;
        CU(cudaMemcpyAsync(&used, c->rec->d_next, 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        const size_t cap = (size_t)c->rec->seg_chunks * c->rec->segs.size();
        if (used <= cap) break;
        // every id below cap was handed out; the failed requests pushed the counter past it
        const unsigned int capu = (unsigned int)cap;
        CU(cudaMemcpyAsync(c->rec->d_next, &capu, 4, cudaMemcpyHostToDevice, c->stream));
        if (rec_add_segment(c->rec)) return 1;
        const unsigned th = 256, bl = (unsigned)((c->B + th - 1) / th);
        mask_outcome_kernel<<<bl, th, 0, c->stream>>>(c->d_outcome, HY_OUTCOME_PAUSED_POOL, c->d_active, c->B);
        CU(cudaGetLastError());
        a.use_active = true;
        a.resume = 1;
        a.rec_append = 1;
        if (launch_once(c, a)) return 1;
    }
    CU(cudaEventRecord(c->ev1, c->stream));
    return 0;
}

// After the stream has been synchronised: kernel time of the call.
void finish_timing(hy_ctx *c)
{
    if (!c->last_launches) return;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->last_ms = ms;
}

int alloc_lanes(hy_ctx *c)
{
    const hy_dims &d = c->d;
    const size_t B = std::max<size_t>(1, c->B), rb = c->rb;
    CU(cudaMalloc(&c->d_state, B * d.n_state * rb));
    CU(cudaMalloc(&c->d_pars, B * std::max<size_t>(1, d.n_par) * rb));
    // one allocation for the per-lane vectors
    //   thi tlo lasth tf tfhi tflo mdt minh maxh (9 x rb) | outcome nsteps (2 x 8) | gidx (4) | active (1) | counter
    const size_t vec = B * rb;
    void *blk = nullptr;
    const size_t bytes = 9 * ((vec + 15) / 16 * 16) + 2 * B * 8 + (B * 4 + 15) / 16 * 16 + (B + 15) / 16 * 16 + 16;
    CU(cudaMalloc(&blk, bytes));
    CU(cudaMemset(blk, 0, bytes));
    char *q = (char *)blk;
    const size_t vs = (vec + 15) / 16 * 16;
    c->d_thi = q, q += vs;
    c->d_tlo = q, q += vs;
    c->d_lasth = q, q += vs;
    c->d_tf = q, q += vs;
    c->d_tfhi = q, q += vs;
    c->d_tflo = q, q += vs;
    c->d_mdt = q, q += vs;
    c->d_minh = q, q += vs;
    c->d_maxh = q, q += vs;
    c->d_outcome = (long long *)q, q += B * 8;
    c->d_nsteps = (unsigned long long *)q, q += B * 8;
    c->d_gidx = (uint32_t *)q, q += (B * 4 + 15) / 16 * 16;
    c->d_active = (unsigned char *)q, q += (B + 15) / 16 * 16;
    c->d_counter = (unsigned int *)q;
    CU(cudaMemset(c->d_state, 0, B * d.n_state * rb));
    CU(cudaMemset(c->d_pars, 0, B * std::max<size_t>(1, d.n_par) * rb));
    if (d.n_events) {
        const size_t ne = d.n_events, nte = std::max<size_t>(1, d.n_tevents);
        CU(cudaMalloc(&c->d_ev_dir, ne * 4));
        CU(cudaMemcpy(c->d_ev_dir, c->h_ev_dir.data(), ne * 4, cudaMemcpyHostToDevice));
        CU(cudaMalloc(&c->d_ev_cd, nte * 8));
        CU(cudaMemcpy(c->d_ev_cd, c->h_ev_cd.data(), nte * 8, cudaMemcpyHostToDevice));
        CU(cudaMalloc(&c->d_cd_elapsed, B * nte * rb));
        CU(cudaMalloc(&c->d_cd_total, B * nte * rb));
        c->log_cap = std::min<unsigned long long>(std::max<unsigned long long>(1ULL << 20, 16ULL * B), 1ULL << 26);
        CU(cudaMalloc(&c->d_log, c->log_cap * sizeof(hy_event_rec)));
        CU(cudaMalloc(&c->d_log_count, 8));
        CU(cudaMemset(c->d_log_count, 0, 8));
    }
    return 0;
}

int common_init(hy_ctx *c)
{
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    return 0;
}

} // namespace

extern "C" {

const char *hy_last_error(void) { return g_err.c_str(); }

int hy_device_count(int *count)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    *count = n;
    return 0;
}

int hy_create2(hy_ctx **out, int device, int fp_bits, const hy_tape *full, const hy_tape *ode, const hy_event_tape *evt,
               const int32_t *ev_dir, const double *ev_cooldown, double tol, int high_accuracy, uint32_t batch)
{
    if (!out || !full || !full->dims || !full->ops || !full->level_start) return fail("hy_create: null argument");
    const hy_dims *dims = full->dims;
    const uint32_t *ev_ref = full->ev_ref;
    if (fp_bits != 32 && fp_bits != 64) return fail("hy_create: fp_bits must be 32 or 64");
    if (dims->order < 2 || dims->order > 62) return fail("hy_create: unsupported Taylor order");
    if (dims->n_events && dims->order + 1 > (uint32_t)hy::EV_MAXP1)
        return fail("hy_create: event detection supports Taylor orders up to 31");
    if (dims->n_tevents > dims->n_events) return fail("hy_create: n_tevents > n_events");
    if (dims->n_events && (!ev_ref || !ev_dir)) return fail("hy_create: null event arrays");
    if ((ode != nullptr) != (evt != nullptr)) return fail("hy_create2: the ODE tape and the event tape come together");
    if (ode && (!ode->dims || !ode->ops || ode->dims->n_events != 0 || ode->dims->n_state != dims->n_state ||
                ode->dims->order != dims->order || evt->n_events != dims->n_events || !evt->ev_ref || !evt->op_start))
        return fail("hy_create2: inconsistent ODE / event tapes");
    int ndev = 0;
    if (hy_device_count(&ndev)) return 1;
    if (ndev == 0) return fail("hy_create: no CUDA device is visible (libhy_cuda has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail("hy_create: invalid device index");
    CU(cudaSetDevice(device));
    hy_ctx *c = new hy_ctx();
    c->device = device;
    c->fp_bits = fp_bits;
    c->rb = fp_bits / 8;
    c->d = *dims;
    c->B = batch;
    c->tol = tol;
    c->high_accuracy = (high_accuracy & HY_CREATE_HIGH_ACCURACY) ? 1 : 0;
    c->no_jit = (high_accuracy & HY_CREATE_COMPACT) ? 1 : 0;
    *out = c;
    const hy_dims &d = c->d;
    if (common_init(c)) return 1;
    c->h_ops.assign(full->ops, full->ops + d.n_ops);
    if (d.n_terms) c->h_terms.assign(full->terms, full->terms + d.n_terms);
    c->h_levels.assign(full->level_start, full->level_start + d.n_levels + 1);
    if (d.n_events) {
        c->h_ev_ref.assign(ev_ref, ev_ref + d.n_events);
        c->h_ev_dir.assign(ev_dir, ev_dir + d.n_events);
        c->h_ev_cd.assign(std::max<size_t>(1, d.n_tevents), -1.0);
        for (size_t i = 0; i < d.n_tevents; ++i) c->h_ev_cd[i] = ev_cooldown ? ev_cooldown[i] : -1.0;
    }
    if (ode && d.n_events) {
        c->d_ode = *ode->dims;
        c->h_ops_ode.assign(ode->ops, ode->ops + c->d_ode.n_ops);
        if (c->d_ode.n_terms) c->h_terms_ode.assign(ode->terms, ode->terms + c->d_ode.n_terms);
        if (evt->n_ops) c->h_evt_ops.assign(evt->ops, evt->ops + evt->n_ops);
        if (evt->n_terms) c->h_evt_terms.assign(evt->terms, evt->terms + evt->n_terms);
        c->h_evt_ref.assign(evt->ev_ref, evt->ev_ref + evt->n_events);
        c->h_evt_start.assign(evt->op_start, evt->op_start + evt->n_events + 1);
        c->evt_rows = evt->n_rows;
        c->have_evt = true;
    }
    if (alloc_lanes(c)) return 1;
    if (d.n_events && hy_reset_cooldowns(c, -1)) return 1;
    if (choose_geometry(c)) return 1;
    return 0;
}

int hy_create(hy_ctx **out, int device, int fp_bits, const hy_dims *dims, const hy_op *ops, const hy_term *terms,
              const uint32_t *level_start, const uint32_t *ev_ref, const int32_t *ev_dir, const double *ev_cooldown,
              double tol, int high_accuracy, uint32_t batch)
{
    if (!dims) return fail("hy_create: null argument");
    hy_tape full{dims, ops, terms, level_start, ev_ref};
    return hy_create2(out, device, fp_bits, &full, nullptr, nullptr, ev_dir, ev_cooldown, tol, high_accuracy, batch);
}

/* Deep copy of a context onto `device` (reference: copy.deepcopy(ta) per ensemble iteration,
 * _ensemble_impl.py:47).  The scheduled program is reused (no matching / scheduling), the lane
 * data (state, pars, time, last_h, results, tc, cooldowns) is copied device-to-device. */
int hy_clone(const hy_ctx *src, hy_ctx **out, int device)
{
    if (!src || !out) return fail("hy_clone: null argument");
    int ndev = 0;
    if (hy_device_count(&ndev)) return 1;
    if (device < 0) device = src->device;
    if (device >= ndev) return fail("hy_clone: invalid device index");
    CU(cudaSetDevice(src->device));
    CU(cudaStreamSynchronize(src->stream));
    CU(cudaSetDevice(device));
    hy_ctx *c = new hy_ctx();
    *out = c;
    c->device = device;
    c->fp_bits = src->fp_bits;
    c->rb = src->rb;
    c->d = src->d;
    c->B = src->B;
    c->tol = src->tol;
    c->high_accuracy = src->high_accuracy;
    c->h_ops = src->h_ops;
    c->h_terms = src->h_terms;
    c->h_levels = src->h_levels;
    c->h_ev_ref = src->h_ev_ref;
    c->h_ev_dir = src->h_ev_dir;
    c->h_ev_cd = src->h_ev_cd;
    c->prog = src->prog;
    c->li = src->li;
    c->TS = src->TS;
    c->d_ode = src->d_ode;
    c->h_ops_ode = src->h_ops_ode;
    c->h_terms_ode = src->h_terms_ode;
    c->h_evt_ops = src->h_evt_ops;
    c->h_evt_terms = src->h_evt_terms;
    c->h_evt_ref = src->h_evt_ref;
    c->h_evt_start = src->h_evt_start;
    c->evt_rows = src->evt_rows;
    c->have_evt = src->have_evt;
    c->use_evt = src->use_evt;
    c->evt_dev = src->evt_dev;
    c->evt_blob = src->evt_blob;
    c->evt_prog = src->evt_prog;
    c->no_jit = src->no_jit;
    c->jit_img = src->jit_img;
    c->jit_img_plain = src->jit_img_plain;
    c->jit_defs = src->jit_defs;
    if (common_init(c)) return 1;
    if (alloc_lanes(c)) return 1;
    {
        // a different device may have another SM count: keep the schedule, refit the grid
        cudaDeviceProp prop{};
        CU(cudaGetDeviceProperties(&prop, device));
        c->li.n_sm = (uint32_t)prop.multiProcessorCount;
        const uint32_t T = c->li.traj_per_cta;
        c->li.ctas = std::max(1u, std::min((c->B + T - 1) / T, c->li.n_sm * env_u32("HY_CUDA_CTAS_PER_SM", 1)));
    }
    if (!c->jit_img.cubin.empty()) {
        const bool whole = c->li.kernel_variant == HY_VARIANT_JIT;
        const std::string lerr = hy::jit::load(c->jit_img, c->li.smem_bytes, whole && !c->li.ws_in_smem, c->jit_k);
        if (!lerr.empty()) return fail("hy_clone: " + lerr);
    }
    if (!c->jit_img_plain.cubin.empty()) {
        const std::string lerr = hy::jit::load(c->jit_img_plain, c->li.smem_bytes, false, c->jit_k_plain);
        if (!lerr.empty()) return fail("hy_clone: " + lerr);
    }
    if (upload_program(c)) return 1;
    const hy_dims &d = c->d;
    const size_t B = c->B, rb = c->rb;
    auto cp = [&](void *dst, const void *s_, size_t bytes) -> cudaError_t {
        if (!bytes || !dst || !s_) return cudaSuccess;
        return cudaMemcpyPeerAsync(dst, device, s_, src->device, bytes, c->stream);
    };
    CU(cp(c->d_state, src->d_state, B * d.n_state * rb));
    CU(cp(c->d_pars, src->d_pars, B * d.n_par * rb));
    CU(cp(c->d_thi, src->d_thi, B * rb));
    CU(cp(c->d_tlo, src->d_tlo, B * rb));
    CU(cp(c->d_lasth, src->d_lasth, B * rb));
    CU(cp(c->d_minh, src->d_minh, B * rb));
    CU(cp(c->d_maxh, src->d_maxh, B * rb));
    CU(cp(c->d_outcome, src->d_outcome, B * 8));
    CU(cp(c->d_nsteps, src->d_nsteps, B * 8));
    if (src->d_tc) {
        if (ensure_tc(c)) return 1;
        CU(cp(c->d_tc, src->d_tc, (size_t)d.n_state * (d.order + 1) * B * rb));
    }
    if (d.n_tevents) {
        CU(cp(c->d_cd_elapsed, src->d_cd_elapsed, B * d.n_tevents * rb));
        CU(cp(c->d_cd_total, src->d_cd_total, B * d.n_tevents * rb));
    }
    if (src->n_red) {
        CU(cudaMalloc((void **)&c->d_red, src->n_red * 4));
        CU(cp(c->d_red, src->d_red, src->n_red * 4));
        c->n_red = src->n_red;
    }
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int hy_destroy(hy_ctx *c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    void *ptrs[] = {c->d_prog, c->d_phase, c->d_ev, c->d_srow, c->d_ssp, c->d_gjet, c->d_state, c->d_pars,
                    c->d_thi /* block of the per-lane vectors */, c->d_tc, c->d_gws, c->d_tmp_in, c->d_tmp_out,
                    c->d_ev_dir, c->d_ev_cd, c->d_cd_elapsed, c->d_cd_total, c->d_log, c->d_log_count, c->d_red, c->d_evt, c->d_evstats,
                    c->d_recoff, c->d_evws};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    hy::jit::unload(c->jit_k);
    hy::jit::unload(c->jit_k_plain);
    rec_free(c->rec);
    rec_free(c->rec_spare);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int hy_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr) return fail("null argument");
    CU(cudaHostAlloc(ptr, bytes ? bytes : 8, cudaHostAllocPortable));
    std::memset(*ptr, 0, bytes ? bytes : 8);
    return 0;
}

int hy_host_free(void *ptr)
{
    if (ptr) CU(cudaFreeHost(ptr));
    return 0;
}

int hy_set_stream(hy_ctx *c, void *cuda_stream)
{
    if (!c) return fail("null ctx");
    CU(cudaSetDevice(c->device));
    if (c->own_stream && c->stream) {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaStreamDestroy(c->stream));
    }
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    return 0;
}

int hy_sync(hy_ctx *c)
{
    if (!c) return fail("null ctx");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

// Stream-ordered: the copies are queued on the context's stream and the call returns (pageable
// sources are staged by the runtime before it returns; pinned sources must stay untouched until
// the next synchronising call - every step/propagate/download call is one).
int hy_upload(hy_ctx *c, const void *state, const void *pars, const void *t_hi, const void *t_lo)
{
    if (!c) return fail("null ctx");
    CU(cudaSetDevice(c->device));
    const size_t B = c->B, rb = c->rb;
    if (B == 0) return 0;
    if (state) CU(cudaMemcpyAsync(c->d_state, state, B * c->d.n_state * rb, cudaMemcpyHostToDevice, c->stream));
    if (pars && c->d.n_par) CU(cudaMemcpyAsync(c->d_pars, pars, B * c->d.n_par * rb, cudaMemcpyHostToDevice, c->stream));
    if (t_hi) CU(cudaMemcpyAsync(c->d_thi, t_hi, B * rb, cudaMemcpyHostToDevice, c->stream));
    if (t_lo) CU(cudaMemcpyAsync(c->d_tlo, t_lo, B * rb, cudaMemcpyHostToDevice, c->stream));
    return 0;
}

int hy_download(hy_ctx *c, void *state, void *t_hi, void *t_lo, void *last_h)
{
    if (!c) return fail("null ctx");
    CU(cudaSetDevice(c->device));
    const size_t B = c->B, rb = c->rb;
    if (B == 0) return 0;
    if (state) CU(cudaMemcpyAsync(state, c->d_state, B * c->d.n_state * rb, cudaMemcpyDeviceToHost, c->stream));
    if (t_hi) CU(cudaMemcpyAsync(t_hi, c->d_thi, B * rb, cudaMemcpyDeviceToHost, c->stream));
    if (t_lo) CU(cudaMemcpyAsync(t_lo, c->d_tlo, B * rb, cudaMemcpyDeviceToHost, c->stream));
    if (last_h) CU(cudaMemcpyAsync(last_h, c->d_lasth, B * rb, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    finish_timing(c);
    return 0;
}

int hy_upload_dev(hy_ctx *c, const void *d_state, const void *d_pars, const void *d_t_hi, const void *d_t_lo)
{
    if (!c) return fail("null ctx");
    CU(cudaSetDevice(c->device));
    const size_t B = c->B, rb = c->rb;
    if (B == 0) return 0;
    if (d_state) CU(cudaMemcpyAsync(c->d_state, d_state, B * c->d.n_state * rb, cudaMemcpyDeviceToDevice, c->stream));
    if (d_pars && c->d.n_par)
        CU(cudaMemcpyAsync(c->d_pars, d_pars, B * c->d.n_par * rb, cudaMemcpyDeviceToDevice, c->stream));
    if (d_t_hi) CU(cudaMemcpyAsync(c->d_thi, d_t_hi, B * rb, cudaMemcpyDeviceToDevice, c->stream));
    if (d_t_lo) CU(cudaMemcpyAsync(c->d_tlo, d_t_lo, B * rb, cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

int hy_state_dev(hy_ctx *c, void **d_state, void **d_t_hi, void **d_t_lo)
{
    if (!c) return fail("null ctx");
    if (d_state) *d_state = c->d_state;
    if (d_t_hi) *d_t_hi = c->d_thi;
    if (d_t_lo) *d_t_lo = c->d_tlo;
    return 0;
}

/* Restore the device-side Taylor coefficients / last step sizes of a copied or unpickled
 * integrator (expose_batch_integrators.cpp:665-669: copies carry tc and last_h). */
int hy_set_tc(hy_ctx *c, const void *tc)
{
    if (!c || !tc) return fail("hy_set_tc: null argument");
    CU(cudaSetDevice(c->device));
    if (ensure_tc(c)) return 1;
    const size_t bytes = (size_t)c->d.n_state * (c->d.order + 1) * c->B * c->rb;
    if (bytes) CU(cudaMemcpyAsync(c->d_tc, tc, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int hy_set_last_h(hy_ctx *c, const void *last_h)
{
    if (!c || !last_h) return fail("hy_set_last_h: null argument");
    CU(cudaSetDevice(c->device));
    if (c->B) CU(cudaMemcpyAsync(c->d_lasth, last_h, c->B * c->rb, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

/* The reference's callback.angle_reducer (expose_callbacks.cpp:67-72) as a device-side post-step
 * op of the propagate_* calls: the listed state variables are reduced to [0, 2 pi) after every
 * step.  n = 0 switches it off. */
int hy_set_angle_reducer(hy_ctx *c, const uint32_t *idx, uint32_t n)
{
    if (!c) return fail("null ctx");
    if (n && !idx) return fail("hy_set_angle_reducer: null index array");
    for (uint32_t i = 0; i < n; ++i)
        if (idx[i] >= c->d.n_state) return fail("hy_set_angle_reducer: state index out of range");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    if (c->d_red) cudaFree(c->d_red);
    c->d_red = nullptr;
    c->n_red = 0;
    if (n) {
        CU(cudaMalloc((void **)&c->d_red, n * 4));
        CU(cudaMemcpy(c->d_red, idx, n * 4, cudaMemcpyHostToDevice));
        c->n_red = n;
    }
    return 0;
}

static int fetch_results(hy_ctx *c, int64_t *outcome, void *min_h, void *max_h, uint64_t *n_steps)
{
    const size_t B = c->B, rb = c->rb;
    if (B == 0) return 0;
    if (outcome) CU(cudaMemcpyAsync(outcome, c->d_outcome, B * 8, cudaMemcpyDeviceToHost, c->stream));
    if (min_h) CU(cudaMemcpyAsync(min_h, c->d_minh, B * rb, cudaMemcpyDeviceToHost, c->stream));
    if (max_h) CU(cudaMemcpyAsync(max_h, c->d_maxh, B * rb, cudaMemcpyDeviceToHost, c->stream));
    if (n_steps) CU(cudaMemcpyAsync(n_steps, c->d_nsteps, B * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    finish_timing(c);
    return 0;
}

int hy_step(hy_ctx *c, const void *max_delta_t, int backward, int write_tc, int64_t *outcome, void *h)
{
    if (!c) return fail("null ctx");
    CU(cudaSetDevice(c->device));
    const size_t B = c->B, rb = c->rb;
    if (max_delta_t && B) CU(cudaMemcpyAsync(c->d_mdt, max_delta_t, B * rb, cudaMemcpyHostToDevice, c->stream));
    RunArgs a;
    a.mode = hy::MODE_STEP;
    a.backward = backward;
    a.max_steps = 1;
    a.have_mdt = max_delta_t != nullptr;
    a.write_tc = write_tc;
    if (run_kernel(c, a)) return 1;
    if (h && B) CU(cudaMemcpyAsync(h, c->d_lasth, B * rb, cudaMemcpyDeviceToHost, c->stream));
    return fetch_results(c, outcome, nullptr, nullptr, nullptr);
}

int hy_propagate_ex(hy_ctx *c, const hy_prop_args *pa, int64_t *outcome, void *min_h, void *max_h, uint64_t *n_steps)
{
    if (!c || !pa) return fail("hy_propagate_ex: null argument");
    const bool grid = pa->grid != nullptr;
    if (!grid && !pa->t && c->B && !pa->resume) return fail("hy_propagate: null time array");
    if (grid && (!pa->grid_out || pa->grid_k == 0)) return fail("hy_propagate_grid: null/empty grid");
    if (grid && pa->c_output) return fail("hy_propagate_grid: no continuous output in grid mode");
    CU(cudaSetDevice(c->device));
    const size_t B = c->B, rb = c->rb, n = c->d.n_state;
    if (B == 0) return 0;
    RunArgs a;
    a.mode = grid ? hy::MODE_GRID : (pa->is_delta ? hy::MODE_FOR : hy::MODE_UNTIL);
    a.max_steps = pa->max_steps;
    a.have_mdt = pa->max_delta_t != nullptr;
    a.write_tc = pa->write_tc || pa->c_output || grid;
    a.resume = pa->resume ? 1 : 0;
    a.pause_on_nt = pa->pause_on_nt;
    a.launch_steps = pa->launch_steps;
    if (pa->max_delta_t) CU(cudaMemcpyAsync(c->d_mdt, pa->max_delta_t, B * rb, cudaMemcpyHostToDevice, c->stream));
    if (pa->active) {
        CU(cudaMemcpyAsync(c->d_active, pa->active, B, cudaMemcpyHostToDevice, c->stream));
        a.use_active = true;
    }
    const unsigned th = 256, bl = (unsigned)((B + th - 1) / th);
    if (grid) {
        const size_t k = pa->grid_k;
        if (ensure_tmp(c, k * B * rb, k * n * B * rb)) return 1;
        if (!a.resume) {
            CU(cudaMemcpyAsync(c->d_tmp_in, pa->grid, k * B * rb, cudaMemcpyHostToDevice, c->stream));
            // NaN-fill: grid points past an early exit stay NaN (reference behaviour).
            CU(cudaMemsetAsync(c->d_tmp_out, 0xff, k * n * B * rb, c->stream));
        }
        a.grid = c->d_tmp_in;
        a.gout = c->d_tmp_out;
        a.grid_k = (uint32_t)k;
    } else if (!a.resume) {
        CU(cudaMemcpyAsync(c->d_tf, pa->t, B * rb, cudaMemcpyHostToDevice, c->stream));
        if (c->fp_bits == 64)
            hy::prep_tf_kernel<double><<<bl, th, 0, c->stream>>>((const double *)c->d_tf, pa->is_delta,
                                                                 (const double *)c->d_thi, (const double *)c->d_tlo,
                                                                 (double *)c->d_tfhi, (double *)c->d_tflo, (uint32_t)B);
        else
            hy::prep_tf_kernel<float><<<bl, th, 0, c->stream>>>((const float *)c->d_tf, pa->is_delta,
                                                                (const float *)c->d_thi, (const float *)c->d_tlo,
                                                                (float *)c->d_tfhi, (float *)c->d_tflo, (uint32_t)B);
        CU(cudaGetLastError());
    }
    if (pa->c_output) {
        const bool append = pa->c_output == 2 && c->rec != nullptr;
        if (!c->rec) {
            if (c->rec_spare) {
                c->rec = c->rec_spare;
                c->rec_spare = nullptr;
            } else if (rec_create(c, &c->rec)) {
                rec_free(c->rec);
                c->rec = nullptr;
                return 1;
            }
        }
        if (!append && rec_reset(c->rec, c->stream)) return 1;
        c->rec->indexed = false;
        a.rec_on = 1;
        a.rec_append = append ? 1 : 0;
    }
    if (run_kernel(c, a)) return 1;
    if (grid) CU(cudaMemcpyAsync(pa->grid_out, c->d_tmp_out, pa->grid_k * n * B * rb, cudaMemcpyDeviceToHost, c->stream));
    return fetch_results(c, outcome, min_h, max_h, n_steps);
}

int hy_propagate(hy_ctx *c, const void *t, int is_delta, uint64_t max_steps, const void *max_delta_t, int write_tc,
                 int c_output, int64_t *outcome, void *min_h, void *max_h, uint64_t *n_steps)
{
    hy_prop_args a{};
    a.t = t;
    a.is_delta = is_delta;
    a.max_steps = max_steps;
    a.max_delta_t = max_delta_t;
    a.write_tc = write_tc;
    a.c_output = c_output ? 1 : 0;
    return hy_propagate_ex(c, &a, outcome, min_h, max_h, n_steps);
}

int hy_propagate_grid(hy_ctx *c, const void *grid, size_t k, uint64_t max_steps, const void *max_delta_t, void *out,
                      int64_t *outcome, void *min_h, void *max_h, uint64_t *n_steps)
{
    if (!grid || !out || k == 0) return fail("hy_propagate_grid: null/empty grid");
    hy_prop_args a{};
    a.max_steps = max_steps;
    a.max_delta_t = max_delta_t;
    a.grid = grid;
    a.grid_k = k;
    a.grid_out = out;
    return hy_propagate_ex(c, &a, outcome, min_h, max_h, n_steps);
}

int hy_last_timing(hy_ctx *c, double *kernel_ms, uint64_t *launches)
{
    if (!c) return fail("null ctx");
    if (kernel_ms) *kernel_ms = c->last_ms;
    if (launches) *launches = c->last_launches;
    return 0;
}

int hy_get_tc(hy_ctx *c, void *tc)
{
    if (!c) return fail("null ctx");
    CU(cudaSetDevice(c->device));
    if (ensure_tc(c)) return 1;
    size_t bytes = (size_t)c->d.n_state * (c->d.order + 1) * c->B * c->rb;
    if (bytes) CU(cudaMemcpyAsync(tc, c->d_tc, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int hy_dense_eval(hy_ctx *c, const void *t, int rel_time, void *out)
{
    if (!c || !t || !out) return fail("hy_dense_eval: null argument");
    CU(cudaSetDevice(c->device));
    const size_t B = c->B, rb = c->rb, n = c->d.n_state;
    if (B == 0) return 0;
    if (ensure_tc(c)) return 1;
    if (ensure_tmp(c, B * rb, n * B * rb)) return 1;
    CU(cudaMemcpyAsync(c->d_tmp_in, t, B * rb, cudaMemcpyHostToDevice, c->stream));
    const unsigned th = 128, bl = (unsigned)((B + th - 1) / th);
    if (c->fp_bits == 64)
        hy::dense_eval_kernel<double><<<bl, th, 0, c->stream>>>(
            (const double *)c->d_tc, (const double *)c->d_thi, (const double *)c->d_tlo, (const double *)c->d_lasth,
            (const double *)c->d_tmp_in, rel_time, (double *)c->d_tmp_out, (uint32_t)n, c->d.order, (uint32_t)B);
    else
        hy::dense_eval_kernel<float><<<bl, th, 0, c->stream>>>(
            (const float *)c->d_tc, (const float *)c->d_thi, (const float *)c->d_tlo, (const float *)c->d_lasth,
            (const float *)c->d_tmp_in, rel_time, (float *)c->d_tmp_out, (uint32_t)n, c->d.order, (uint32_t)B);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, c->d_tmp_out, n * B * rb, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

/* ---- continuous output ---- */
int hy_cout_detach(hy_ctx *c, hy_cout **out)
{
    if (!c || !out) return fail("hy_cout_detach: null argument");
    *out = nullptr;
    if (!c->rec) return 0;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    *out = c->rec;
    c->rec = nullptr;
    return 0;
}

extern "C++" {
namespace {
// host image of a record pool: per lane a run of consecutive chunks, records in the tc layout
template <typename R>
void build_pool_image(std::vector<R> &buf, std::vector<uint32_t> &head, std::vector<uint32_t> &tail, uint32_t n, uint32_t P1,
                      uint32_t B, const uint64_t *ns, const R *tcs, const R *thi, const R *tlo, uint32_t rec_len,
                      uint32_t chunk_len)
{
    const uint32_t nP = n * P1;
    uint32_t cid = 0;
    for (uint32_t l = 0; l < B; ++l) {
        const uint32_t cnt = (uint32_t)ns[l], nch = (cnt + hy::HY_REC_CH - 1u) / hy::HY_REC_CH;
        head[l] = nch ? cid : hy::HY_REC_NONE;
        tail[l] = nch ? cid + nch - 1u : hy::HY_REC_NONE;
        for (uint32_t j = 0; j < nch; ++j) {
            const uint32_t next = j + 1u < nch ? cid + j + 1u : hy::HY_REC_NONE;
            std::memcpy(&buf[(size_t)(cid + j) * chunk_len], &next, 4);
        }
        for (uint32_t s = 0; s < cnt; ++s) {
            R *dst = &buf[(size_t)(cid + s / hy::HY_REC_CH) * chunk_len + 2u + (s % hy::HY_REC_CH) * rec_len];
            for (uint32_t i = 0; i < nP; ++i) dst[i] = tcs[((size_t)s * nP + i) * B + l];
            dst[nP] = thi[(size_t)(s + 1) * B + l];
            dst[nP + 1] = tlo[(size_t)(s + 1) * B + l];
        }
        cid += nch;
    }
}
} // namespace
} // extern "C++"

int hy_cout_from_host(hy_cout **out, int device, int fp_bits, uint32_t n_state, uint32_t order, uint32_t batch,
                      const uint64_t *n_steps, const void *tcs, const void *times_hi, const void *times_lo, uint64_t S)
{
    if (!out || !n_steps || !times_hi || !times_lo || (S && !tcs)) return fail("hy_cout_from_host: null argument");
    if (fp_bits != 32 && fp_bits != 64) return fail("hy_cout_from_host: fp_bits must be 32 or 64");
    if (!batch || !n_state) return fail("hy_cout_from_host: empty record");
    *out = nullptr;
    size_t total = 0;
    for (uint32_t l = 0; l < batch; ++l) {
        if (n_steps[l] > S) return fail("hy_cout_from_host: a lane holds more steps than the arrays");
        total += (n_steps[l] + hy::HY_REC_CH - 1u) / hy::HY_REC_CH;
    }
    if (total >= 0x7fffffffu) return fail("hy_cout_from_host: the record is too large");
    CU(cudaSetDevice(device));
    hy_cout *r = new hy_cout();
    r->device = device;
    r->rb = fp_bits == 64 ? 8 : 4;
    r->B = batch;
    r->n = n_state;
    r->P1 = order + 1;
    r->rec_len = r->n * r->P1 + 2u;
    r->chunk_len = 2u + hy::HY_REC_CH * r->rec_len;
    r->si = r->P1; // the reference's tc layout
    r->sk = 1;
    r->seg_chunks = (uint32_t)std::max<size_t>(16, total);
    auto bail = [&](int rc) {
        rec_free(r);
        return rc;
    };
    if (cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail("hy_cout_from_host: no stream"));
    if (cudaMalloc((void **)&r->d_seg, HY_REC_MAXSEG * sizeof(void *)) != cudaSuccess ||
        cudaMemset(r->d_seg, 0, HY_REC_MAXSEG * sizeof(void *)) != cudaSuccess)
        return bail(fail("hy_cout_from_host: out of device memory"));
    const size_t B = r->B;
    const size_t lane_bytes = 16 + 5 * B * 4 + 2 * B * 8;
    if (cudaMalloc(&r->d_lane, lane_bytes) != cudaSuccess || cudaMemset(r->d_lane, 0, lane_bytes) != cudaSuccess)
        return bail(fail("hy_cout_from_host: out of device memory"));
    char *q = (char *)r->d_lane;
    r->d_next = (unsigned int *)q;
    r->d_dir_next = (unsigned int *)(q + 8);
    q += 16;
    r->d_t0hi = q, q += B * 8;
    r->d_t0lo = q, q += B * 8;
    r->d_head = (uint32_t *)q, q += B * 4;
    r->d_tail = (uint32_t *)q, q += B * 4;
    r->d_count = (uint32_t *)q, q += B * 4;
    r->d_nch = (uint32_t *)q, q += B * 4;
    r->d_dir_off = (uint32_t *)q;
    if (rec_add_segment(r)) return bail(1);
    std::vector<uint32_t> head(B), tail(B), count(B);
    for (size_t l = 0; l < B; ++l) count[l] = (uint32_t)n_steps[l];
    const size_t elems = (size_t)r->seg_chunks * r->chunk_len;
    cudaError_t e = cudaSuccess;
    if (r->rb == 8) {
        std::vector<double> buf(elems, 0.0);
        build_pool_image<double>(buf, head, tail, r->n, r->P1, r->B, n_steps, (const double *)tcs, (const double *)times_hi,
                                 (const double *)times_lo, r->rec_len, r->chunk_len);
        e = cudaMemcpy(r->segs[0], buf.data(), elems * 8, cudaMemcpyHostToDevice);
    } else {
        std::vector<float> buf(elems, 0.0f);
        build_pool_image<float>(buf, head, tail, r->n, r->P1, r->B, n_steps, (const float *)tcs, (const float *)times_hi,
                                (const float *)times_lo, r->rec_len, r->chunk_len);
        e = cudaMemcpy(r->segs[0], buf.data(), elems * 4, cudaMemcpyHostToDevice);
    }
    const unsigned int used = (unsigned int)total;
    if (e == cudaSuccess) e = cudaMemcpy(r->d_next, &used, 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(r->d_t0hi, times_hi, B * r->rb, cudaMemcpyHostToDevice); // row 0: the start times
    if (e == cudaSuccess) e = cudaMemcpy(r->d_t0lo, times_lo, B * r->rb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(r->d_head, head.data(), B * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(r->d_tail, tail.data(), B * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(r->d_count, count.data(), B * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return bail(fail(std::string("hy_cout_from_host: ") + cudaGetErrorString(e)));
    r->indexed = false;
    *out = r;
    return 0;
}

int hy_cout_free(hy_cout *r, hy_ctx *recycle_into)
{
    if (!r) return 0;
    hy_ctx *c = recycle_into;
    if (c && !c->rec_spare && c->device == r->device && c->rb == r->rb && c->B == r->B && c->d.n_state == r->n &&
        c->d.order + 1 == r->P1) {
        c->rec_spare = r; // the next recording of this context reuses the pool: no cudaMalloc in steady state
        return 0;
    }
    rec_free(r);
    return 0;
}

int hy_cout_info(hy_cout *r, uint64_t *n_steps, uint64_t *max_steps)
{
    if (!r) return fail("hy_cout_info: null record");
    CU(cudaSetDevice(r->device));
    std::vector<uint32_t> cnt(r->B);
    if (r->B) CU(cudaMemcpy(cnt.data(), r->d_count, (size_t)r->B * 4, cudaMemcpyDeviceToHost));
    uint64_t mx = 0;
    for (uint32_t l = 0; l < r->B; ++l) {
        if (n_steps) n_steps[l] = cnt[l];
        mx = std::max<uint64_t>(mx, cnt[l]);
    }
    if (max_steps) *max_steps = mx;
    return 0;
}

int hy_cout_get(hy_cout *r, void *tcs, void *times_hi, void *times_lo, uint64_t S)
{
    if (!r) return fail("hy_cout_get: null record");
    CU(cudaSetDevice(r->device));
    if (rec_index(r)) return 1;
    const size_t B = r->B, rb = r->rb, nP = (size_t)r->n * r->P1;
    const size_t tcs_bytes = tcs ? S * nP * B * rb : 0, tm_bytes = (S + 1) * B * rb;
    if (rec_ensure_tmp(r, 2 * tm_bytes, std::max<size_t>(8, tcs_bytes))) return 1;
    const unsigned th = 128;
    const unsigned bl = (unsigned)(((S + 1) * B + th - 1) / th);
    char *thi = (char *)r->d_tmp_in, *tlo = thi + tm_bytes;
    if (rb == 8)
        hy::rec_gather_kernel<double><<<bl, th, 0, r->stream>>>(rec_dev<double>(r, 1, 0), r->d_dir_off, r->d_dir,
                                                                tcs ? (double *)r->d_tmp_out : nullptr, (double *)thi,
                                                                (double *)tlo, r->n, r->P1 - 1, r->B, (uint32_t)S);
    else
        hy::rec_gather_kernel<float><<<bl, th, 0, r->stream>>>(rec_dev<float>(r, 1, 0), r->d_dir_off, r->d_dir,
                                                               tcs ? (float *)r->d_tmp_out : nullptr, (float *)thi,
                                                               (float *)tlo, r->n, r->P1 - 1, r->B, (uint32_t)S);
    CU(cudaGetLastError());
    if (tcs && tcs_bytes) CU(cudaMemcpyAsync(tcs, r->d_tmp_out, tcs_bytes, cudaMemcpyDeviceToHost, r->stream));
    if (times_hi) CU(cudaMemcpyAsync(times_hi, thi, tm_bytes, cudaMemcpyDeviceToHost, r->stream));
    if (times_lo) CU(cudaMemcpyAsync(times_lo, tlo, tm_bytes, cudaMemcpyDeviceToHost, r->stream));
    CU(cudaStreamSynchronize(r->stream));
    return 0;
}

int hy_cout_eval(hy_cout *r, const void *t, size_t k, void *out)
{
    if (!r || !t || !out) return fail("hy_cout_eval: null argument");
    CU(cudaSetDevice(r->device));
    const size_t B = r->B, rb = r->rb, n = r->n;
    if (B == 0 || k == 0) return 0;
    if (rec_index(r)) return 1;
    if (rec_ensure_tmp(r, k * B * rb, k * n * B * rb)) return 1;
    CU(cudaMemcpyAsync(r->d_tmp_in, t, k * B * rb, cudaMemcpyHostToDevice, r->stream));
    const unsigned th = 128;
    const unsigned bl = (unsigned)((k * B + th - 1) / th);
    if (rb == 8)
        hy::cout_eval_kernel<double><<<bl, th, 0, r->stream>>>(rec_dev<double>(r, 1, 0), r->d_dir_off, r->d_dir,
                                                               (const double *)r->d_tmp_in, (double *)r->d_tmp_out,
                                                               (uint32_t)n, r->P1 - 1, (uint32_t)B, (uint32_t)k);
    else
        hy::cout_eval_kernel<float><<<bl, th, 0, r->stream>>>(rec_dev<float>(r, 1, 0), r->d_dir_off, r->d_dir,
                                                              (const float *)r->d_tmp_in, (float *)r->d_tmp_out,
                                                              (uint32_t)n, r->P1 - 1, (uint32_t)B, (uint32_t)k);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, r->d_tmp_out, k * n * B * rb, cudaMemcpyDeviceToHost, r->stream));
    CU(cudaStreamSynchronize(r->stream));
    return 0;
}

/* hy_cout_eval with device-resident query times / output (same layouts): the evaluation of a
 * continuous output whose consumers live on the GPU. */
int hy_cout_eval_dev(hy_cout *r, const void *d_t, size_t k, void *d_out)
{
    if (!r || !d_t || !d_out) return fail("hy_cout_eval_dev: null argument");
    CU(cudaSetDevice(r->device));
    const size_t B = r->B, rb = r->rb, n = r->n;
    if (B == 0 || k == 0) return 0;
    if (rec_index(r)) return 1;
    const unsigned th = 128;
    const unsigned bl = (unsigned)((k * B + th - 1) / th);
    if (rb == 8)
        hy::cout_eval_kernel<double><<<bl, th, 0, r->stream>>>(rec_dev<double>(r, 1, 0), r->d_dir_off, r->d_dir,
                                                               (const double *)d_t, (double *)d_out, (uint32_t)n, r->P1 - 1,
                                                               (uint32_t)B, (uint32_t)k);
    else
        hy::cout_eval_kernel<float><<<bl, th, 0, r->stream>>>(rec_dev<float>(r, 1, 0), r->d_dir_off, r->d_dir,
                                                              (const float *)d_t, (float *)d_out, (uint32_t)n, r->P1 - 1,
                                                              (uint32_t)B, (uint32_t)k);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(r->stream));
    return 0;
}

int hy_events_count(hy_ctx *c, uint64_t *n)
{
    if (!c || !n) return fail("null argument");
    *n = 0;
    if (!c->d_log_count) return 0;
    CU(cudaSetDevice(c->device));
    unsigned long long v = 0;
    CU(cudaMemcpyAsync(&v, c->d_log_count, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (v > c->log_cap)
        return fail("the device event log overflowed (" + std::to_string(v) + " events, capacity " +
                    std::to_string(c->log_cap) + "): propagate in shorter segments");
    *n = v;
    return 0;
}

int hy_events_drain(hy_ctx *c, hy_event_rec *recs, uint64_t cap, uint64_t *n)
{
    if (!c || !n) return fail("null argument");
    *n = 0;
    if (!c->d_log_count) return 0;
    uint64_t have = 0;
    if (hy_events_count(c, &have)) return 1;
    if (have > cap)
        return fail("hy_events_drain: " + std::to_string(have) + " events are logged but the buffer holds " +
                    std::to_string(cap) + " (nothing was dropped: call again with a larger buffer)");
    if (have && recs) CU(cudaMemcpyAsync(recs, c->d_log, have * sizeof(hy_event_rec), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemsetAsync(c->d_log_count, 0, 8, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *n = have;
    return 0;
}

int hy_get_cooldowns(hy_ctx *c, void *elapsed, void *total)
{
    if (!c) return fail("null ctx");
    if (!c->d.n_tevents) return 0;
    CU(cudaSetDevice(c->device));
    const size_t bytes = (size_t)c->B * c->d.n_tevents * c->rb;
    if (elapsed) CU(cudaMemcpyAsync(elapsed, c->d_cd_elapsed, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (total) CU(cudaMemcpyAsync(total, c->d_cd_total, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int hy_set_cooldowns(hy_ctx *c, const void *elapsed, const void *total)
{
    if (!c) return fail("null ctx");
    if (!c->d.n_tevents) return 0;
    CU(cudaSetDevice(c->device));
    const size_t bytes = (size_t)c->B * c->d.n_tevents * c->rb;
    if (elapsed) CU(cudaMemcpyAsync(c->d_cd_elapsed, elapsed, bytes, cudaMemcpyHostToDevice, c->stream));
    if (total) CU(cudaMemcpyAsync(c->d_cd_total, total, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int hy_reset_cooldowns(hy_ctx *c, int64_t lane)
{
    if (!c) return fail("null ctx");
    if (!c->d.n_tevents || !c->d_cd_total) return 0;
    CU(cudaSetDevice(c->device));
    const size_t nte = c->d.n_tevents, B = c->B;
    if (lane >= (int64_t)B) return fail("hy_reset_cooldowns: lane out of range");
    const size_t first = lane < 0 ? 0 : (size_t)lane, count = lane < 0 ? B : 1;
    // total = -1 (not in cooldown), elapsed = 0
    if (c->fp_bits == 64) {
        std::vector<double> m1(count * nte, -1.0);
        CU(cudaMemcpyAsync((double *)c->d_cd_total + first * nte, m1.data(), m1.size() * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    } else {
        std::vector<float> m1(count * nte, -1.0f);
        CU(cudaMemcpyAsync((float *)c->d_cd_total + first * nte, m1.data(), m1.size() * 4, cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    CU(cudaMemsetAsync((char *)c->d_cd_elapsed + first * nte * c->rb, 0, count * nte * c->rb, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int hy_get_launch_info(hy_ctx *c, hy_launch_info *info)
{
    if (!c || !info) return fail("null argument");
    *info = c->li;
    return 0;
}

/* Debug counters of the event path of the register-resident kernels (HY_CUDA_EVENT_STATS=1 at
 * hy_create time): steps taken, and steps whose interval enclosure could not exclude an event. */
int hy_get_event_stats(hy_ctx *c, uint64_t *steps, uint64_t *full_evals)
{
    if (!c) return fail("null ctx");
    unsigned long long v[2] = {0, 0};
    if (c->d_evstats) {
        CU(cudaSetDevice(c->device));
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaMemcpy(v, c->d_evstats, 16, cudaMemcpyDeviceToHost));
    }
    if (steps) *steps = v[0];
    if (full_evals) *full_evals = v[1];
    return 0;
}

int hy_get_device(hy_ctx *c, int *device)
{
    if (!c || !device) return fail("null argument");
    *device = c->device;
    return 0;
}

int hy_tape_kernel_variant(const hy_dims *dims, const hy_op *ops, const hy_term *terms, uint32_t *variant)
{
    if (!dims || !ops || !variant) return fail("hy_tape_kernel_variant: null argument");
    hy::NbMatch m;
    *variant = 0;
    // (the introspection call does not know the precision: FP64 assumed for orders above 20)
    if (hy::match_nbody(*dims, ops, terms, m)) *variant = hy::nbody_kernel_variant(m.nb, dims->order, 64);
    hy::CrbMatch cm; // (FP64 order, then FP32 order: the introspection call does not know the precision)
    if (!*variant && (hy::match_cr3bp(*dims, ops, terms, 64, cm) || hy::match_cr3bp(*dims, ops, terms, 32, cm)))
        *variant = hy::cr3bp_kernel_variant(dims->order, 64);
    return 0;
}

/* Generate and compile the kernel of a tape without a device (fills the on-disk cache; the build
 * step runs it for the BASELINE tapes so that the GPU box starts warm). */
int hy_jit_precompile(int fp_bits, const hy_tape *full, uint32_t batch, int *from_cache, double *compile_s)
{
    if (!full || !full->dims || !full->ops) return fail("hy_jit_precompile: null argument");
    if (fp_bits != 32 && fp_bits != 64) return fail("hy_jit_precompile: fp_bits must be 32 or 64");
    hy::Program pr;
    hy::jit::Image img;
    bool smem = false;
    uint32_t T = 0;
    std::string err;
    // (sm_100a: 148 SMs, 227 KB of shared memory per CTA)
    const int r = jit_plan(*full->dims, full->ops, full->terms, full->ev_ref, fp_bits, batch, 148u, 232448u,
                           env_u32("HY_CUDA_JIT", 2) == 1, pr, smem, T, img, err);
    if (r < 0) return fail("hy_jit_precompile: " + err);
    if (from_cache) *from_cache = r == 0 ? -1 : (img.from_cache ? 1 : 0);
    if (compile_s) *compile_s = img.compile_s;
    return 0;
}

/* The same for the event functions of a system served by a register-resident kernel: match the ODE
 * tape, lower the event tape, generate and compile the kernel that hy_create2 would build for it
 * (*from_cache = -1: the system gets no such kernel - its events run on the interpreter). */
int hy_jit_precompile_events(int fp_bits, const hy_tape *full, const hy_tape *ode, const hy_event_tape *evt, int *from_cache,
                             double *compile_s)
{
    if (!full || !full->dims || !ode || !ode->dims || !ode->ops || !evt || !evt->ops)
        return fail("hy_jit_precompile_events: null argument");
    if (fp_bits != 32 && fp_bits != 64) return fail("hy_jit_precompile_events: fp_bits must be 32 or 64");
    const hy_dims &d = *full->dims;
    if (from_cache) *from_cache = -1;
    if (compile_s) *compile_s = 0;
    uint32_t variant = 0, G = 0;
    std::vector<uint32_t> state_row;
    std::string defs;
    hy::NbMatch nbm;
    hy::CrbMatch crm;
    if (hy::match_nbody(*ode->dims, ode->ops, ode->terms, nbm) && (variant = hy::nbody_kernel_variant(nbm.nb, d.order, fp_bits))) {
        G = nbm.g32 ? 32u : 16u;
        for (uint32_t i = 0; i < d.n_state; ++i) state_row.push_back((uint32_t)hy::nbr_state_off((int)i));
        defs = std::string(nbm.has_par ? "#define HY_NBR_PAR 1\n" : "") + (nbm.g32 ? "#define HY_NBR_G32 1\n" : "");
    } else if (hy::match_cr3bp(*ode->dims, ode->ops, ode->terms, fp_bits, crm)) {
        variant = hy::cr3bp_kernel_variant(d.order, fp_bits);
        G = 2;
        for (uint32_t i = 0; i < d.n_state; ++i) state_row.push_back(i);
    }
    const std::string kname = evt_kernel_name(variant, fp_bits);
    if (kname.empty()) return 0;
    hy::EvtProgram ep;
    const std::vector<hy_op> ops(evt->ops, evt->ops + evt->n_ops);
    const std::vector<hy_term> terms(evt->terms, evt->terms + evt->n_terms);
    const std::vector<uint32_t> ev_ref(evt->ev_ref, evt->ev_ref + evt->n_events);
    const std::vector<uint32_t> op_start(evt->op_start, evt->op_start + evt->n_events + 1);
    const std::string err = hy::build_event_program(d.n_state, d.order, ops, terms, ev_ref, op_start, evt->n_rows, ep);
    if (!err.empty()) return 0; // (hy_create2 keeps such a system on the interpreter)
    const std::string src = hy::jit::evt_kernel_source(ep, state_row, d.order, defs, G, 2u * (d.n_state + ep.n_slots));
    hy::jit::Image img;
    const std::string jerr = hy::jit::build(src, kname, img);
    if (!jerr.empty()) return fail("hy_jit_precompile_events: " + jerr);
    if (from_cache) *from_cache = img.from_cache ? 1 : 0;
    if (compile_s) *compile_s = img.compile_s;
    return 0;
}

int hy_measure_fma_peak(int device, int fp_bits, double *tflops)
{
    CU(cudaSetDevice(device));
    cudaDeviceProp prop{};
    CU(cudaGetDeviceProperties(&prop, device));
    const int threads = 512, blocks = prop.multiProcessorCount * 4, iters = 4096;
    void *out = nullptr;
    CU(cudaMalloc(&out, (size_t)threads * blocks * 8));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(e0));
        if (fp_bits == 64)
            hy::fma_peak_kernel<double><<<blocks, threads>>>((double *)out, iters);
        else
            hy::fma_peak_kernel<float><<<blocks, threads>>>((float *)out, iters);
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 8 * 16 * (double)iters * threads * blocks;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    CU(cudaFree(out));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return 0;
}

} // extern "C"
