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

struct hy_ctx {
    int device = 0;
    int fp_bits = 64;
    size_t rb = 8; // bytes per real
    hy_dims d{};
    uint32_t B = 0;
    double tol = 0;
    int high_accuracy = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // tape (device)
    std::vector<hy_op> h_ops; // the ABI tape (kept for re-scheduling)
    std::vector<hy_term> h_terms;
    std::vector<uint32_t> h_levels;
    hy::Program prog;
    void *d_prog = nullptr; // [ops | terms | imm]
    uint32_t *d_phase = nullptr;
    uint32_t *d_ev = nullptr;       // remapped event jet rows (device layout)
    std::vector<uint32_t> h_ev_ref; // ABI event references
    uint32_t *d_srow = nullptr;
    int32_t *d_ssp = nullptr;
    void *d_gjet = nullptr;
    // lanes (device)
    void *d_state = nullptr, *d_pars = nullptr, *d_thi = nullptr, *d_tlo = nullptr, *d_lasth = nullptr;
    void *d_tf = nullptr, *d_mdt = nullptr, *d_minh = nullptr, *d_maxh = nullptr, *d_tc = nullptr;
    long long *d_outcome = nullptr;
    unsigned long long *d_nsteps = nullptr;
    unsigned int *d_counter = nullptr;
    void *d_gws = nullptr;
    // events
    int32_t *d_ev_dir = nullptr;
    double *d_ev_cd = nullptr;
    void *d_cd_elapsed = nullptr, *d_cd_total = nullptr;
    hy_event_rec *d_log = nullptr;
    unsigned long long *d_log_count = nullptr;
    unsigned long long log_cap = 0;
    // continuous output
    void *d_cout_tcs = nullptr, *d_cout_thi = nullptr, *d_cout_tlo = nullptr;
    unsigned long long *d_cout_count = nullptr;
    uint64_t cout_S = 0; // recorded capacity == max steps over lanes
    // scratch for grid / dense / cout evaluation
    void *d_tmp_in = nullptr, *d_tmp_out = nullptr;
    size_t tmp_in_bytes = 0, tmp_out_bytes = 0;
    // launch geometry
    hy_launch_info li{};
    uint32_t TS = 0;
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

template <typename R> cudaError_t launch(const hy::KParams<R> &P, const hy_launch_info &li, cudaStream_t s)
{
    // register-resident N-body kernels (hy_nbody_reg.cuh)
    switch (li.kernel_variant) { // instantiated in hy_nb3.cu ... hy_nb6.cu
    case 0: break;
    case 3: return hy::launch_nbody_kernel<R, 3>(P, li, s);
    case 4: return hy::launch_nbody_kernel<R, 4>(P, li, s);
    case 5: return hy::launch_nbody_kernel<R, 5>(P, li, s);
    case 6: return hy::launch_nbody_kernel<R, 6>(P, li, s);
    case hy::NBR_VARIANT_P22: // 6 bodies, unrolled to order 22 (hy_nb6.cu)
        if constexpr (std::is_same<R, double>::value) return hy::launch_nbody_kernel_p22(P, li, s);
        return cudaErrorInvalidValue;
    case 106: // warpgroup rotation (experimental)
        if constexpr (std::is_same<R, double>::value) return hy::launch_nbody_kernel_wgx(P, li, s);
        return cudaErrorInvalidValue;
    case hy::CRB_VARIANT: return hy::launch_cr3bp_kernel<R>(P, li, s); // hy_cr3bp.cu
    case hy::CRB_VARIANT_P22:
        if constexpr (std::is_same<R, double>::value) return hy::launch_cr3bp_kernel_p22(P, li, s);
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
    return hy::ProgDims{p.n_slots, p.n_tslots, (uint32_t)p.imm.size(), p.n_phases,
                        p.ws_len,  p.par_off,  p.one_off,              p.n_spill};
}

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
    hy::NbMatch nbm;
    if (!force_global && !Genv && env_u32("HY_CUDA_NO_NBODY_REG", 0) == 0 &&
        hy::match_nbody(d, c->h_ops.data(), c->h_terms.data(), nbm) &&
        hy::nbody_kernel_variant(nbm.nb, d.order, c->fp_bits)) {
        hy::Program pr;
        pr.G = 16;
        pr.n_phases = 0;
        pr.phase_slot = {0};
        pr.imm = nbm.imm;
        // experimental: warpgroup rotation (24 trajectories per SM, registers traded between warpgroups)
        const bool wgx = env_u32("HY_CUDA_WGX", 0) != 0 && nbm.nb == 6 && c->fp_bits == 64 &&
                         d.order == (uint32_t)hy::NBR_PMAX && c->B >= 24u * li.n_sm;
        pr.ws_len = (uint32_t)hy::nbr_ws(wgx);
        pr.par_off = pr.one_off = pr.ws_len;
        pr.n_spill = 0;
        for (uint32_t i = 0; i < d.n_state; ++i) pr.state_row.push_back((uint32_t)hy::nbr_state_off((int)i));
        pr.state_spill.assign(d.n_state, -1);
        pr.n_clusters = nbm.n_pairs;
        pr.lane_utilisation = (double)nbm.n_pairs / 16.0;
        // column stride: even (16-byte aligned vectors); + 2 spreads the two trajectories of a warp over the banks
        const uint32_t RS = (pr.ws_len + 1u) / 2u * 2u + 2u;
        hy::SmemLayout L0 = hy::make_layout(d, prog_dims(pr), 16, 0, RS, (uint32_t)c->rb, 0);
        const uint32_t fixed = L0.total + 64;
        if (fixed < (uint32_t)smem_optin && ((uint32_t)smem_optin - fixed) / (RS * (uint32_t)c->rb) >= 2) {
            bestG = 16;
            bestT = std::min(((uint32_t)smem_optin - fixed) / (RS * (uint32_t)c->rb), (wgx ? 384u : max_threads) / 16u) & ~1u;
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
    // Register-resident CR3BP kernel (hy_cr3bp_reg.cuh): two lanes per trajectory, state jets
    // [order][variable] in shared memory.
    hy::CrbMatch crm;
    if (!li.kernel_variant && !force_global && !Genv && env_u32("HY_CUDA_NO_CR3BP_REG", 0) == 0 &&
        hy::match_cr3bp(d, c->h_ops.data(), c->h_terms.data(), c->fp_bits, crm)) {
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
        uint32_t RS = pr.ws_len;
        while (RS % 4u != 2u) ++RS;
        hy::SmemLayout L0 = hy::make_layout(d, prog_dims(pr), 2, 0, RS, (uint32_t)c->rb, 0);
        const uint32_t fixed = L0.total + 64;
        if (fixed < (uint32_t)smem_optin && ((uint32_t)smem_optin - fixed) / (RS * (uint32_t)c->rb) >= 16) {
            bestG = 2;
            const uint32_t mt = (uint32_t)hy::hy_max_threads(2, true, -1, (int)c->rb);
            bestT = std::min(((uint32_t)smem_optin - fixed) / (RS * (uint32_t)c->rb), mt / 2u) & ~15u;
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
        uint32_t Tfit = budget / (RS * (uint32_t)c->rb);
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
    if (!bestG) return fail("hy_create: the program does not fit in shared memory for any group size");
    uint32_t G = bestG, T = bestT;
    li.ws_in_smem = best_smem ? 1 : 0;
    uint32_t Tenv = env_u32("HY_CUDA_TRAJ_PER_CTA", 0);
    if (Tenv) T = std::min(Tenv, T);
    // Do not keep more trajectories resident than the batch can feed.
    uint32_t per_cta_needed = std::max(1u, (c->B + li.n_sm - 1) / li.n_sm);
    T = std::max(1u, std::min(T, per_cta_needed));
    if (li.kernel_variant) T = (T + 1u) & ~1u; // whole warps: the two trajectories of a warp step in lockstep
    if (li.kernel_variant == 106) T = 24;      // whole warpgroups
    if (li.kernel_variant == (uint32_t)hy::CRB_VARIANT || li.kernel_variant == (uint32_t)hy::CRB_VARIANT_P22)
        T = (T + 15u) & ~15u; // whole warps (16 trajectories)
    li.group = G;
    li.traj_per_cta = T;
    li.threads = ((T * G + 31) / 32) * 32;
    uint32_t ctas = (c->B + T - 1) / T;
    li.ctas = std::max(1u, std::min(ctas, li.n_sm * env_u32("HY_CUDA_CTAS_PER_SM", 1)));
    const uint32_t RS = bestRS;
    c->TS = RS;
    c->prog = best;
    hy::SmemLayout L = hy::make_layout(d, prog_dims(c->prog), G, T, RS, (uint32_t)c->rb, (int)li.ws_in_smem);
    li.smem_bytes = L.total;
    if (li.smem_bytes > (uint32_t)smem_optin) return fail("tape does not fit in shared memory");
    li.regs_per_thread = (uint32_t)(c->fp_bits == 64 ? regs_for_group<double>(G, li.ws_in_smem, li.kernel_variant)
                                                     : regs_for_group<float>(G, li.ws_in_smem, li.kernel_variant));
    // Upload the program blob [ops | terms | imm] (same layout as in shared memory).
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
    return 0;
}

template <typename R>
hy::KParams<R> make_params(hy_ctx *c, int mode, int backward, uint64_t max_steps, bool have_mdt, int write_tc)
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
    P.tf = (const R *)c->d_tf;
    P.mdt = have_mdt ? (const R *)c->d_mdt : nullptr;
    P.outcome = c->d_outcome;
    P.min_h = (R *)c->d_minh;
    P.max_h = (R *)c->d_maxh;
    P.n_steps = c->d_nsteps;
    P.tc = (R *)c->d_tc;
    P.cout_tcs = nullptr;
    P.cout_thi = nullptr;
    P.cout_tlo = nullptr;
    P.cout_cap = 0;
    P.grid = nullptr;
    P.gout = nullptr;
    P.grid_k = 0;
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
    P.max_steps = max_steps;
    P.mode = mode;
    P.backward = backward;
    P.write_tc = write_tc;
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

struct RunExtras {
    bool record_cout = false;
    const void *grid = nullptr;
    void *gout = nullptr;
    uint32_t grid_k = 0;
};

int ensure_tmp(hy_ctx *c, size_t in_bytes, size_t out_bytes)
{
    if (in_bytes > c->tmp_in_bytes) {
        if (c->d_tmp_in) cudaFree(c->d_tmp_in);
        c->d_tmp_in = nullptr;
        CU(cudaMalloc(&c->d_tmp_in, in_bytes));
        c->tmp_in_bytes = in_bytes;
    }
    if (out_bytes > c->tmp_out_bytes) {
        if (c->d_tmp_out) cudaFree(c->d_tmp_out);
        c->d_tmp_out = nullptr;
        CU(cudaMalloc(&c->d_tmp_out, out_bytes));
        c->tmp_out_bytes = out_bytes;
    }
    return 0;
}

template <typename R> void apply_extras(hy_ctx *c, hy::KParams<R> &P, const RunExtras &x)
{
    if (x.record_cout) {
        P.cout_tcs = (R *)c->d_cout_tcs;
        P.cout_thi = (R *)c->d_cout_thi;
        P.cout_tlo = (R *)c->d_cout_tlo;
        P.cout_cap = (uint32_t)c->cout_S;
    }
    P.grid = (const R *)x.grid;
    P.gout = (R *)x.gout;
    P.grid_k = x.grid_k;
}

int run_kernel(hy_ctx *c, int mode, int backward, uint64_t max_steps, bool have_mdt, int write_tc,
               const RunExtras &x = RunExtras())
{
    if (c->B == 0) {
        c->last_ms = 0;
        c->last_launches = 0;
        return 0;
    }
    if (write_tc && ensure_tc(c)) return 1;
    CU(cudaMemsetAsync(c->d_counter, 0, sizeof(unsigned int), c->stream));
    CU(cudaEventRecord(c->ev0, c->stream));
    cudaError_t e;
    if (c->fp_bits == 64) {
        auto P = make_params<double>(c, mode, backward, max_steps, have_mdt, write_tc);
        apply_extras(c, P, x);
        e = launch<double>(P, c->li, c->stream);
    } else {
        auto P = make_params<float>(c, mode, backward, max_steps, have_mdt, write_tc);
        apply_extras(c, P, x);
        e = launch<float>(P, c->li, c->stream);
    }
    if (e != cudaSuccess) return fail(std::string("kernel launch: ") + cudaGetErrorString(e));
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->last_ms = ms;
    c->last_launches = 1;
    return 0;
}

} // namespace

template <typename R> static void cout_transpose(const hy_ctx *c, const std::vector<R> &src,
                                                 const std::vector<unsigned long long> &cnt, R *dst, uint64_t S)
{
    // device layout [S][B][n*P1] -> reference layout [S][n][P1][B], NaN past each lane's count
    const size_t B = c->B, nP = (size_t)c->d.n_state * (c->d.order + 1);
    for (uint64_t s = 0; s < S; ++s)
        for (size_t l = 0; l < B; ++l) {
            const bool ok = s < cnt[l];
            const R *p = src.data() + (s * B + l) * nP;
            for (size_t i = 0; i < nP; ++i) dst[(s * nP + i) * B + l] = ok ? p[i] : (R)NAN;
        }
}

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

int hy_create(hy_ctx **out, int device, int fp_bits, const hy_dims *dims, const hy_op *ops, const hy_term *terms,
              const uint32_t *level_start, const uint32_t *ev_ref, const int32_t *ev_dir, const double *ev_cooldown,
              double tol, int high_accuracy, uint32_t batch)
{
    if (!out || !dims || !ops || !level_start) return fail("hy_create: null argument");
    if (fp_bits != 32 && fp_bits != 64) return fail("hy_create: fp_bits must be 32 or 64");
    if (dims->order < 2 || dims->order > 62) return fail("hy_create: unsupported Taylor order");
    if (dims->n_events && dims->order + 1 > (uint32_t)hy::EV_MAXP1)
        return fail("hy_create: event detection supports Taylor orders up to 31");
    if (dims->n_tevents > dims->n_events) return fail("hy_create: n_tevents > n_events");
    if (dims->n_events && (!ev_ref || !ev_dir)) return fail("hy_create: null event arrays");
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
    c->high_accuracy = high_accuracy;
    *out = c;
    const hy_dims &d = c->d;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    c->h_ops.assign(ops, ops + d.n_ops);
    if (d.n_terms) c->h_terms.assign(terms, terms + d.n_terms);
    c->h_levels.assign(level_start, level_start + d.n_levels + 1);
    if (d.n_events) c->h_ev_ref.assign(ev_ref, ev_ref + d.n_events);
    const size_t B = std::max<size_t>(1, batch), rb = c->rb;
    CU(cudaMalloc(&c->d_state, B * d.n_state * rb));
    CU(cudaMalloc(&c->d_pars, B * std::max<size_t>(1, d.n_par) * rb));
    CU(cudaMalloc(&c->d_thi, B * rb));
    CU(cudaMalloc(&c->d_tlo, B * rb));
    CU(cudaMalloc(&c->d_lasth, B * rb));
    CU(cudaMalloc(&c->d_tf, B * rb));
    CU(cudaMalloc(&c->d_mdt, B * rb));
    CU(cudaMalloc(&c->d_minh, B * rb));
    CU(cudaMalloc(&c->d_maxh, B * rb));
    CU(cudaMalloc(&c->d_outcome, B * sizeof(long long)));
    CU(cudaMalloc(&c->d_nsteps, B * sizeof(unsigned long long)));
    CU(cudaMalloc(&c->d_counter, sizeof(unsigned int)));
    CU(cudaMemset(c->d_state, 0, B * d.n_state * rb));
    CU(cudaMemset(c->d_pars, 0, B * std::max<size_t>(1, d.n_par) * rb));
    CU(cudaMemset(c->d_thi, 0, B * rb));
    CU(cudaMemset(c->d_tlo, 0, B * rb));
    CU(cudaMemset(c->d_lasth, 0, B * rb));
    if (d.n_events) {
        const size_t ne = d.n_events, nte = std::max<size_t>(1, d.n_tevents);
        CU(cudaMalloc(&c->d_ev_dir, ne * 4));
        CU(cudaMemcpy(c->d_ev_dir, ev_dir, ne * 4, cudaMemcpyHostToDevice));
        std::vector<double> cd(nte, -1.0);
        for (size_t i = 0; i < d.n_tevents; ++i) cd[i] = ev_cooldown ? ev_cooldown[i] : -1.0;
        CU(cudaMalloc(&c->d_ev_cd, nte * 8));
        CU(cudaMemcpy(c->d_ev_cd, cd.data(), nte * 8, cudaMemcpyHostToDevice));
        CU(cudaMalloc(&c->d_cd_elapsed, B * nte * rb));
        CU(cudaMalloc(&c->d_cd_total, B * nte * rb));
        c->log_cap = std::min<unsigned long long>(std::max<unsigned long long>(1ULL << 20, 16ULL * B), 1ULL << 26);
        CU(cudaMalloc(&c->d_log, c->log_cap * sizeof(hy_event_rec)));
        CU(cudaMalloc(&c->d_log_count, 8));
        CU(cudaMemset(c->d_log_count, 0, 8));
        if (hy_reset_cooldowns(c, -1)) return 1;
    }
    if (choose_geometry(c)) return 1;
    return 0;
}

int hy_destroy(hy_ctx *c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    void *ptrs[] = {c->d_prog, c->d_phase, c->d_ev, c->d_srow, c->d_ssp, c->d_gjet,   c->d_state,   c->d_pars,   c->d_thi,     c->d_tlo,
                    c->d_lasth, c->d_tf,   c->d_mdt,    c->d_minh, c->d_maxh,    c->d_tc,     c->d_outcome, c->d_nsteps,
                    c->d_counter, c->d_gws, c->d_cout_tcs, c->d_cout_thi, c->d_cout_tlo, c->d_cout_count,
                    c->d_tmp_in, c->d_tmp_out, c->d_ev_dir, c->d_ev_cd, c->d_cd_elapsed, c->d_cd_total, c->d_log,
                    c->d_log_count};
    for (void *p : ptrs)
        if (p) cudaFree(p);
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
    CU(cudaStreamSynchronize(c->stream));
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
    CU(cudaStreamSynchronize(c->stream));
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

static int fetch_results(hy_ctx *c, int64_t *outcome, void *min_h, void *max_h, uint64_t *n_steps)
{
    const size_t B = c->B, rb = c->rb;
    if (B == 0) return 0;
    if (outcome) CU(cudaMemcpyAsync(outcome, c->d_outcome, B * 8, cudaMemcpyDeviceToHost, c->stream));
    if (min_h) CU(cudaMemcpyAsync(min_h, c->d_minh, B * rb, cudaMemcpyDeviceToHost, c->stream));
    if (max_h) CU(cudaMemcpyAsync(max_h, c->d_maxh, B * rb, cudaMemcpyDeviceToHost, c->stream));
    if (n_steps) CU(cudaMemcpyAsync(n_steps, c->d_nsteps, B * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int hy_step(hy_ctx *c, const void *max_delta_t, int backward, int write_tc, int64_t *outcome, void *h)
{
    if (!c) return fail("null ctx");
    CU(cudaSetDevice(c->device));
    const size_t B = c->B, rb = c->rb;
    if (max_delta_t && B) CU(cudaMemcpyAsync(c->d_mdt, max_delta_t, B * rb, cudaMemcpyHostToDevice, c->stream));
    if (run_kernel(c, hy::MODE_STEP, backward, 1, max_delta_t != nullptr, write_tc)) return 1;
    if (fetch_results(c, outcome, nullptr, nullptr, nullptr)) return 1;
    if (h && B) {
        CU(cudaMemcpyAsync(h, c->d_lasth, B * rb, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int hy_propagate(hy_ctx *c, const void *t, int is_delta, uint64_t max_steps, const void *max_delta_t, int write_tc,
                 int c_output, int64_t *outcome, void *min_h, void *max_h, uint64_t *n_steps)
{
    if (!c) return fail("null ctx");
    if (!t && c->B) return fail("hy_propagate: null time array");
    CU(cudaSetDevice(c->device));
    const size_t B = c->B, rb = c->rb;
    if (B) CU(cudaMemcpyAsync(c->d_tf, t, B * rb, cudaMemcpyHostToDevice, c->stream));
    if (max_delta_t && B) CU(cudaMemcpyAsync(c->d_mdt, max_delta_t, B * rb, cudaMemcpyHostToDevice, c->stream));
    const int mode = is_delta ? hy::MODE_FOR : hy::MODE_UNTIL;
    if (!c_output || B == 0) {
        if (run_kernel(c, mode, 0, max_steps, max_delta_t != nullptr, write_tc)) return 1;
        return fetch_results(c, outcome, min_h, max_h, n_steps);
    }
    // Continuous output.  The number of steps is not known in advance and the
    // stepping is deterministic, so: pass 1 counts the steps on a backup of the
    // state, pass 2 (on the restored state) records exactly that many.
    const size_t n = c->d.n_state, P1 = c->d.order + 1;
    void *bk_state = nullptr, *bk_thi = nullptr, *bk_tlo = nullptr;
    CU(cudaMalloc(&bk_state, B * n * rb));
    CU(cudaMalloc(&bk_thi, B * rb));
    CU(cudaMalloc(&bk_tlo, B * rb));
    CU(cudaMemcpyAsync(bk_state, c->d_state, B * n * rb, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(bk_thi, c->d_thi, B * rb, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(bk_tlo, c->d_tlo, B * rb, cudaMemcpyDeviceToDevice, c->stream));
    if (run_kernel(c, mode, 0, max_steps, max_delta_t != nullptr, 0)) return 1;
    double ms1 = c->last_ms;
    std::vector<unsigned long long> ns(B);
    CU(cudaMemcpy(ns.data(), c->d_nsteps, B * 8, cudaMemcpyDeviceToHost));
    uint64_t S = 0;
    for (auto v : ns) S = std::max<uint64_t>(S, v);
    for (void *p : {c->d_cout_tcs, c->d_cout_thi, c->d_cout_tlo})
        if (p) cudaFree(p);
    c->d_cout_tcs = c->d_cout_thi = c->d_cout_tlo = nullptr;
    if (!c->d_cout_count) CU(cudaMalloc(&c->d_cout_count, B * 8));
    c->cout_S = S;
    const size_t tcs_bytes = std::max<size_t>(8, S * B * n * P1 * rb), tm_bytes = (S + 1) * B * rb;
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    if (tcs_bytes + 2 * tm_bytes > free_b)
        return fail("hy_propagate: continuous output needs " + std::to_string((tcs_bytes + 2 * tm_bytes) >> 20) +
                    " MiB of device memory but only " + std::to_string(free_b >> 20) + " MiB are free");
    CU(cudaMalloc(&c->d_cout_tcs, tcs_bytes));
    CU(cudaMalloc(&c->d_cout_thi, tm_bytes));
    CU(cudaMalloc(&c->d_cout_tlo, tm_bytes));
    CU(cudaMemcpyAsync(c->d_state, bk_state, B * n * rb, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_thi, bk_thi, B * rb, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_tlo, bk_tlo, B * rb, cudaMemcpyDeviceToDevice, c->stream));
    RunExtras x;
    x.record_cout = true;
    if (run_kernel(c, mode, 0, max_steps, max_delta_t != nullptr, 1, x)) return 1;
    c->last_ms += ms1;
    c->last_launches = 2;
    CU(cudaMemcpyAsync(c->d_cout_count, c->d_nsteps, B * 8, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(bk_state);
    cudaFree(bk_thi);
    cudaFree(bk_tlo);
    return fetch_results(c, outcome, min_h, max_h, n_steps);
}

int hy_propagate_grid(hy_ctx *c, const void *grid, size_t k, uint64_t max_steps, const void *max_delta_t, void *out,
                      int64_t *outcome, void *min_h, void *max_h, uint64_t *n_steps)
{
    if (!c) return fail("null ctx");
    if (!grid || !out || k == 0) return fail("hy_propagate_grid: null/empty grid");
    CU(cudaSetDevice(c->device));
    const size_t B = c->B, rb = c->rb, n = c->d.n_state;
    if (B == 0) return 0;
    if (ensure_tmp(c, k * B * rb, k * n * B * rb)) return 1;
    CU(cudaMemcpyAsync(c->d_tmp_in, grid, k * B * rb, cudaMemcpyHostToDevice, c->stream));
    // NaN-fill: grid points past an early exit stay NaN (reference behaviour).
    CU(cudaMemsetAsync(c->d_tmp_out, 0xff, k * n * B * rb, c->stream));
    if (max_delta_t) CU(cudaMemcpyAsync(c->d_mdt, max_delta_t, B * rb, cudaMemcpyHostToDevice, c->stream));
    RunExtras x;
    x.grid = c->d_tmp_in;
    x.gout = c->d_tmp_out;
    x.grid_k = (uint32_t)k;
    if (run_kernel(c, hy::MODE_GRID, 0, max_steps, max_delta_t != nullptr, 1, x)) return 1;
    CU(cudaMemcpyAsync(out, c->d_tmp_out, k * n * B * rb, cudaMemcpyDeviceToHost, c->stream));
    return fetch_results(c, outcome, min_h, max_h, n_steps);
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

int hy_cout_info(hy_ctx *c, uint64_t *n_steps, uint64_t *max_steps)
{
    if (!c) return fail("null ctx");
    CU(cudaSetDevice(c->device));
    if (max_steps) *max_steps = c->d_cout_tcs ? c->cout_S : 0;
    if (n_steps && c->B) {
        if (!c->d_cout_count) {
            std::memset(n_steps, 0, c->B * 8);
        } else {
            CU(cudaMemcpy(n_steps, c->d_cout_count, c->B * 8, cudaMemcpyDeviceToHost));
        }
    }
    return 0;
}

int hy_cout_get(hy_ctx *c, void *tcs, void *times_hi, void *times_lo, uint64_t S)
{
    if (!c) return fail("null ctx");
    if (!c->d_cout_tcs) return fail("hy_cout_get: no continuous output was recorded");
    if (S != c->cout_S) return fail("hy_cout_get: S does not match the recorded number of steps");
    CU(cudaSetDevice(c->device));
    const size_t B = c->B, rb = c->rb, nP = (size_t)c->d.n_state * (c->d.order + 1);
    std::vector<unsigned long long> cnt(B);
    CU(cudaMemcpy(cnt.data(), c->d_cout_count, B * 8, cudaMemcpyDeviceToHost));
    if (tcs && S) {
        if (c->fp_bits == 64) {
            std::vector<double> tmp(S * B * nP);
            CU(cudaMemcpy(tmp.data(), c->d_cout_tcs, tmp.size() * rb, cudaMemcpyDeviceToHost));
            cout_transpose<double>(c, tmp, cnt, (double *)tcs, S);
        } else {
            std::vector<float> tmp(S * B * nP);
            CU(cudaMemcpy(tmp.data(), c->d_cout_tcs, tmp.size() * rb, cudaMemcpyDeviceToHost));
            cout_transpose<float>(c, tmp, cnt, (float *)tcs, S);
        }
    }
    for (int which = 0; which < 2; ++which) {
        void *dst = which ? times_lo : times_hi;
        if (!dst) continue;
        CU(cudaMemcpy(dst, which ? c->d_cout_tlo : c->d_cout_thi, (S + 1) * B * rb, cudaMemcpyDeviceToHost));
        for (uint64_t s = 0; s <= S; ++s)
            for (size_t l = 0; l < B; ++l)
                if (s > cnt[l]) {
                    if (c->fp_bits == 64)
                        ((double *)dst)[s * B + l] = NAN;
                    else
                        ((float *)dst)[s * B + l] = NAN;
                }
    }
    return 0;
}

int hy_cout_eval(hy_ctx *c, const void *t, size_t k, void *out)
{
    if (!c || !t || !out) return fail("hy_cout_eval: null argument");
    if (!c->d_cout_tcs) return fail("hy_cout_eval: no continuous output was recorded");
    CU(cudaSetDevice(c->device));
    const size_t B = c->B, rb = c->rb, n = c->d.n_state;
    if (B == 0 || k == 0) return 0;
    if (ensure_tmp(c, k * B * rb, k * n * B * rb)) return 1;
    CU(cudaMemcpyAsync(c->d_tmp_in, t, k * B * rb, cudaMemcpyHostToDevice, c->stream));
    const unsigned th = 128;
    const unsigned bl = (unsigned)((k * B + th - 1) / th);
    if (c->fp_bits == 64)
        hy::cout_eval_kernel<double><<<bl, th, 0, c->stream>>>(
            (const double *)c->d_cout_tcs, (const double *)c->d_cout_thi, (const double *)c->d_cout_tlo,
            c->d_cout_count, (const double *)c->d_tmp_in, (double *)c->d_tmp_out, (uint32_t)n, c->d.order,
            (uint32_t)B, (uint32_t)k);
    else
        hy::cout_eval_kernel<float><<<bl, th, 0, c->stream>>>(
            (const float *)c->d_cout_tcs, (const float *)c->d_cout_thi, (const float *)c->d_cout_tlo,
            c->d_cout_count, (const float *)c->d_tmp_in, (float *)c->d_tmp_out, (uint32_t)n, c->d.order, (uint32_t)B,
            (uint32_t)k);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, c->d_tmp_out, k * n * B * rb, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int hy_events_count(hy_ctx *c, uint64_t *n)
{
    if (!c || !n) return fail("null argument");
    *n = 0;
    if (!c->d_log_count) return 0;
    CU(cudaSetDevice(c->device));
    unsigned long long v = 0;
    CU(cudaMemcpy(&v, c->d_log_count, 8, cudaMemcpyDeviceToHost));
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
    const uint64_t m = std::min<uint64_t>(have, cap);
    if (m && recs) CU(cudaMemcpy(recs, c->d_log, m * sizeof(hy_event_rec), cudaMemcpyDeviceToHost));
    CU(cudaMemset(c->d_log_count, 0, 8));
    *n = m;
    return 0;
}

int hy_get_cooldowns(hy_ctx *c, void *elapsed, void *total)
{
    if (!c) return fail("null ctx");
    if (!c->d.n_tevents) return 0;
    CU(cudaSetDevice(c->device));
    const size_t bytes = (size_t)c->B * c->d.n_tevents * c->rb;
    if (elapsed) CU(cudaMemcpy(elapsed, c->d_cd_elapsed, bytes, cudaMemcpyDeviceToHost));
    if (total) CU(cudaMemcpy(total, c->d_cd_total, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int hy_set_cooldowns(hy_ctx *c, const void *elapsed, const void *total)
{
    if (!c) return fail("null ctx");
    if (!c->d.n_tevents) return 0;
    CU(cudaSetDevice(c->device));
    const size_t bytes = (size_t)c->B * c->d.n_tevents * c->rb;
    if (elapsed) CU(cudaMemcpy(c->d_cd_elapsed, elapsed, bytes, cudaMemcpyHostToDevice));
    if (total) CU(cudaMemcpy(c->d_cd_total, total, bytes, cudaMemcpyHostToDevice));
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
        CU(cudaMemcpy((double *)c->d_cd_total + first * nte, m1.data(), m1.size() * 8, cudaMemcpyHostToDevice));
    } else {
        std::vector<float> m1(count * nte, -1.0f);
        CU(cudaMemcpy((float *)c->d_cd_total + first * nte, m1.data(), m1.size() * 4, cudaMemcpyHostToDevice));
    }
    CU(cudaMemset((char *)c->d_cd_elapsed + first * nte * c->rb, 0, count * nte * c->rb));
    return 0;
}

int hy_get_launch_info(hy_ctx *c, hy_launch_info *info)
{
    if (!c || !info) return fail("null argument");
    *info = c->li;
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
