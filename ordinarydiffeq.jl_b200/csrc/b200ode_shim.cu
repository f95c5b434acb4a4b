// b200ode_shim.cu — the C ABI of include/b200ode.h.
//
// Host side of the ensemble hot path: NVRTC-compiles the user's RHS/Jacobian C
// source together with the stepper kernels (device/*.cuh, embedded at build time
// in device_sources.inc), loads the cubin through the CUDA runtime's library API,
// lays the ensemble out in HBM and launches b200_initdt + b200_integrate.
// There is no CPU fallback anywhere in this file: without a device every solving
// entry point fails with B200ODE_ECUDA.
#include "../../include/b200ode.h"

#include <cuda_runtime.h>
#include <nvrtc.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <mutex>
#include <thread>
#include <vector>

#include "device_sources.inc"   // k_b200_header_names[], k_b200_header_sources[], k_b200_num_headers
#include "device/b200_base.cuh"  // the flagged fast division / square root, for b200ode_selftest_fastmath

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

int nslots_typed(const B200Problem* prob, const B200Opts* o, int dtype);

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return fail(B200ODE_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));         \
    } while (0)

// Mirror of struct B200Params in device/b200_ensemble.cuh (same member order and types).
template <typename R>
struct Params {
    long long N;
    const R* u0; long long u0_ts, u0_cs;
    const R* p; long long p_ts, p_cs;
    R t0, tf;
    R reltol, abstol;
    R dt_user;
    R dtmin, dtmax;
    long long maxiters;
    const R* saveat;
    int nsaveat;
    int save_start, save_end;
    int nslots;
    R* dt0;
    R* u_final; long long uf_ts, uf_cs;
    R* t_final;
    R* us;
    int* naccept; int* nreject; int* nf; int* retcode; int* nsaved;
    int* njacs; int* nw; int* nsolve;
    unsigned long long* work_counter;
    int flags;
    int tol_const;
    R tol100_tf;
    const long long* row_offsets;
    R* ts_rag;
    R* dts_rag;
    const R* tstops;
    int ntstops;
    R fpe0, rfpe0;
    const R* disc; int ndisc;
    const double* tspans; int dtmax_default;
    R* peer_out[8]; int npeer, peer_world, peer_rank; long long peer_block;
};

struct DevBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (ptr) { cudaFree(ptr); ptr = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e != cudaSuccess) { e = cudaMalloc(&ptr, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
};

}  // namespace

static const unsigned kCounterRing = 256;
struct b200ode_handle_s {
    int device = 0;
    std::atomic<unsigned> launch_seq{0};
    int num_sms = 0;
    cudaStream_t stream = nullptr;       // compute + H2D
    cudaStream_t copy_stream = nullptr;  // D2H of finished chunks
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    std::vector<cudaEvent_t> chunk_events;
    // scratch owned by the handle (grow-only)
    DevBuf counter, dt0, saveat, scratch_t;
    std::vector<double> saveat_cached;   // grid currently resident in `saveat` ...
    int saveat_cached_dtype = -1;        // ... in this real type
    std::vector<double> tstops_cached;   // same for the tstops list
    int tstops_cached_dtype = -1;
    DevBuf in_u0, in_p, out_uf, out_tf, out_us, out_i32, red_partial, stat_out, row_offsets, rag_dts, dense_tq, dense_out, scan_tiles, tstops, in_tspans;
    // pinned bounce buffers for large D2H copies into pageable caller memory (d2h_large)
    void* stage[2] = {nullptr, nullptr};
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
};

struct b200ode_program_s {
    b200ode_handle h = nullptr;
    int alg = 0, dtype = 0, n = 0, np = 0;
    std::vector<char> cubin;
    cudaLibrary_t lib = nullptr;
    cudaKernel_t k_integrate = nullptr, k_initdt = nullptr, k_dense = nullptr;
    int coop_l = 0;              // > 0: lane-group kernel (device/b200_coop.cuh) with this many lanes per trajectory
    bool everystep = false;      // compiled with -DB200_EVERYSTEP=1 (ragged save_everystep output)
    bool tstops = false;         // compiled with -DB200_TSTOPS=1
    bool adaptive = true;        // false: compiled with -DB200_ADAPTIVE=0 (fixed dt)
    bool callbacks = false;      // compiled with a CallbackSet (b200ode_compile_callbacks)
    bool vector_tol = false;     // compiled with -DB200_VECTOR_TOL=1 (per-component abstol / reltol)
    bool tspans = false;         // compiled with -DB200_TSPANS=1 (per-trajectory time spans)
    bool reverse = false;        // compiled with -DB200_REVERSE=1 (tf < t0): the kernels run in mirrored time
    std::vector<double> tolv_cached;   // the 2n tolerances currently resident in the module's B200_TOLV
    int nsave = 0;               // components per saved row: n, or the length of -DB200_SAVE_IDXS=...
    size_t dyn_smem = 0;
    int wide_nt = 0;             // > 0: shared-memory stage kernel (device/b200_vern7_wide.cuh): trajectories in flight per CTA
    // Launches without a saveat grid run a second build of the same program with the saveat handling compiled out
    // (-DB200_NO_SAVEAT=1: 3.21 -> 3.08 ms per 2^20 Lorenz trajectories, final states only).  It is built on first use
    // from the sources kept here; `nosave_ok` is false for program kinds that do not qualify.
    bool nosave_ok = false, nosave_tried = false;
    b200ode_program_s* nosave = nullptr;
    std::mutex nosave_mu;
    std::string keep_rhs, keep_rhs_name, keep_jac, keep_jac_name, keep_tg, keep_tg_name, keep_extra;
    B200ProgramInfo info{};
};

namespace {

std::string strip_includes(const char* src) {
    std::string out;
    if (!src) return out;
    const char* s = src;
    while (*s) {
        const char* e = strchr(s, '\n');
        size_t len = e ? (size_t)(e - s) : strlen(s);
        std::string line(s, len);
        size_t i = line.find_first_not_of(" \t");
        bool is_inc = (i != std::string::npos && line.compare(i, 1, "#") == 0 &&
                       line.find("include", i) != std::string::npos);
        if (!is_inc) { out += line; }
        out += '\n';
        if (!e) break;
        s = e + 1;
    }
    return out;
}

bool is_identifier(const char* s) {
    if (!s || !*s) return false;
    if (!(isalpha((unsigned char)s[0]) || s[0] == '_')) return false;
    for (const char* c = s; *c; ++c)
        if (!(isalnum((unsigned char)*c) || *c == '_')) return false;
    return true;
}

bool is_stiff_alg(int alg) {      // Rosenbrock-type: need jac + tgrad, report njacs/nw/nsolve
    return alg == B200ODE_ALG_ROSENBROCK23 || alg == B200ODE_ALG_ROSENBROCK32 || alg == B200ODE_ALG_RODAS5P ||
           alg == B200ODE_ALG_RODAS5PE || alg == B200ODE_ALG_AUTOTSIT5_ROSENBROCK23 || alg == B200ODE_ALG_RODAS3P || alg == B200ODE_ALG_RODAS23W ||
           (alg >= B200ODE_ALG_RODAS5 && alg <= B200ODE_ALG_RODAS4P2);
}

// alg_order of the reference (alg_utils.jl of each solver package)
int alg_order(int alg) {
    switch (alg) {
        case B200ODE_ALG_TSIT5: case B200ODE_ALG_DP5: case B200ODE_ALG_RODAS5P: case B200ODE_ALG_RODAS5PE: case B200ODE_ALG_RODAS5: return 5;
        case B200ODE_ALG_AUTOTSIT5_ROSENBROCK23: return 5;      // the branch a trajectory starts in (Tsit5)
        case B200ODE_ALG_VERN6: return 6;
        case B200ODE_ALG_VERN7: return 7;
        case B200ODE_ALG_VERN8: return 8;
        case B200ODE_ALG_VERN9: return 9;
        case B200ODE_ALG_ROSENBROCK23: return 2;
        case B200ODE_ALG_BS3: case B200ODE_ALG_ROSENBROCK32: case B200ODE_ALG_RODAS3P: case B200ODE_ALG_RODAS23W: return 3;
        default: return 4;      // Rodas4, Rodas42, Rodas4P, Rodas4P2
    }
}

// -DB200_SAVE_IDXS=i0,i1,...  -> number of listed components (0: option absent, -1: malformed / out of range)
int parse_save_idxs(const char* extra_options, int n) {
    const char* key = "-DB200_SAVE_IDXS=";
    const char* at = extra_options ? strstr(extra_options, key) : nullptr;
    if (!at) return 0;
    at += strlen(key);
    int count = 0;
    while (*at && *at != ' ') {
        if (!isdigit((unsigned char)*at)) return -1;
        long v = 0;
        while (isdigit((unsigned char)*at)) { v = v * 10 + (*at - '0'); if (v > 100000) return -1; ++at; }
        if (v >= n) return -1;
        ++count;
        if (*at == ',') { ++at; if (!isdigit((unsigned char)*at)) return -1; }
        else if (*at && *at != ' ') return -1;
    }
    return count > 0 ? count : -1;
}

int validate_compile_args(int alg, int dtype, int n, int np, const char* rhs_src, const char* rhs_name,
                          const char* jac_src, const char* jac_name, const char* tgrad_src,
                          const char* tgrad_name) {
    if (alg < B200ODE_ALG_TSIT5 || alg > B200ODE_ALG_RODAS23W)
        return fail(B200ODE_EINVAL, "alg must be one of the B200ODE_ALG_* constants");
    if (dtype != B200ODE_F64 && dtype != B200ODE_F32) return fail(B200ODE_EINVAL, "dtype must be B200ODE_F64 or B200ODE_F32");
    if (n < 1 || n > 64) return fail(B200ODE_EINVAL, "state dimension n must be in 1..64 (one trajectory per thread)");
    if (np < 0 || np > 256) return fail(B200ODE_EINVAL, "parameter dimension np must be in 0..256");
    if (!rhs_src || !is_identifier(rhs_name)) return fail(B200ODE_EINVAL, "rhs_src and a valid rhs_name are required");
    bool stiff = is_stiff_alg(alg);
    if (stiff) {
        if (!jac_src || !is_identifier(jac_name))
            return fail(B200ODE_EINVAL, "Rosenbrock algorithms need jac_src/jac_name (ODEFunction(f; jac, tgrad))");
        if (tgrad_src && !is_identifier(tgrad_name)) return fail(B200ODE_EINVAL, "tgrad_name is not an identifier");
    }
    return B200ODE_OK;
}

// The CallbackSet of a program: forward declarations + the user's condition / affect sources, the parameter table
// B200_CB_TABLE and the two dispatchers device/b200_callbacks.cuh calls.  Continuous callbacks come first, then the
// discrete ones, each group in the caller's order (CallbackSet(continuous..., discrete...)).
int callbacks_source(int alg, int dtype, const B200CallbackSrc* cbs, int ncb, bool everystep, bool coop, std::string& out) {
    if (!cbs || ncb < 1 || ncb > 16) return fail(B200ODE_EINVAL, "callbacks: between 1 and 16 callbacks");
    if (coop) return fail(B200ODE_EUNSUPPORTED, "callbacks / isoutofdomain are not available in the lane-group kernel");
    std::vector<int> order;
    int isout_at = -1;
    for (int pass = 1; pass >= 0; --pass)
        for (int i = 0; i < ncb; ++i) {
            if (cbs[i].kind != B200ODE_CB_DISCRETE && cbs[i].kind != B200ODE_CB_CONTINUOUS && cbs[i].kind != B200ODE_CB_ISOUTOFDOMAIN)
                return fail(B200ODE_EINVAL, "callbacks: kind must be one of the B200ODE_CB_* constants");
            if (cbs[i].kind == pass) order.push_back(i);
            if (cbs[i].kind == B200ODE_CB_ISOUTOFDOMAIN && pass == 0) {
                if (isout_at >= 0) return fail(B200ODE_EINVAL, "callbacks: at most one isoutofdomain function");
                isout_at = i;
            }
        }
    int ncc = 0;
    for (int i = 0; i < ncb; ++i) ncc += (cbs[i].kind == B200ODE_CB_CONTINUOUS);
    if (ncc > 0 && alg != B200ODE_ALG_TSIT5)
        return fail(B200ODE_EUNSUPPORTED, "continuous callbacks are available for Tsit5 (discrete callbacks and isoutofdomain: every stepper)");
    std::vector<std::string> emitted;
    auto add_fn = [&](const char* src, const char* name, bool is_condition) -> int {
        if (!is_identifier(name)) return fail(B200ODE_EINVAL, "callbacks: a function name is not an identifier");
        for (auto& e : emitted) if (e == name) return B200ODE_OK;          // one definition may serve several callbacks
        if (!src) return fail(B200ODE_EINVAL, std::string("callbacks: no source for ") + name);
        emitted.push_back(name);
        if (is_condition) out += std::string("__device__ __forceinline__ real ") + name + "(const real* u, const real* p, const real t);\n";
        else out += std::string("__device__ __forceinline__ void ") + name + "(real* u, real* p, const real t, int* terminate);\n";
        out += strip_includes(src);
        out += "\n";
        return B200ODE_OK;
    };
    out += "// ---- user source (callbacks) ----\n";
    if (isout_at >= 0) {        // the isoutofdomain keyword (solve.jl:166; integrator_utils.jl:612): any algorithm
        int rc = add_fn(cbs[isout_at].condition_src, cbs[isout_at].condition_name, true);
        if (rc) return rc;
        out += std::string("#define B200_ISOUT(u, p, t) ") + cbs[isout_at].condition_name + "((u), (p), (t))\n";
    }
    if (order.empty()) return B200ODE_OK;
    ncb = (int)order.size();
    char buf[512];
    std::string table = "#define B200_CB_TABLE { ", cond = "", aff = "";
    for (size_t k = 0; k < order.size(); ++k) {
        const B200CallbackSrc& c = cbs[order[k]];
        const bool cont = c.kind == B200ODE_CB_CONTINUOUS;
        int rc = add_fn(c.condition_src, c.condition_name, true);
        if (rc) return rc;
        const bool has_aff = c.affect_name != nullptr, has_neg = cont && c.affect_neg_name != nullptr;
        if (has_aff && (rc = add_fn(c.affect_src, c.affect_name, false))) return rc;
        if (has_neg && (rc = add_fn(c.affect_neg_src, c.affect_neg_name, false))) return rc;
        if (!everystep && (c.save_before || c.save_after))
            return fail(B200ODE_EUNSUPPORTED, "callbacks with save_positions need the ragged output: compile with B200ODE_OPT_EVERYSTEP "
                                              "(and pass B200ODE_FLAG_NO_STEP_ROWS for save_everystep = false), or use save_positions = (false, false)");
        if (c.rootfind < -1 || c.rootfind > 2) return fail(B200ODE_EINVAL, "callbacks: rootfind must be -1 (default), 0, 1 or 2");
        const int rootfind = c.rootfind < 0 ? 1 : c.rootfind;
        const int ip = c.interp_points < 0 ? 10 : c.interp_points;
        const double eps = dtype == B200ODE_F32 ? 1.1920928955078125e-07 : 2.220446049250313e-16;
        // abstol = 10eps (of Float64 in SciMLBase's constructor; converted to the real type when it is compared)
        const double abstol = c.abstol < 0 ? 10.0 * 2.220446049250313e-16 : c.abstol;
        (void)eps;
        const double nudge = c.repeat_nudge < 0 ? 0.01 : c.repeat_nudge;
        snprintf(buf, sizeof buf, "{%d, %d, %d, %d, %d, %d, %d, %.17g, (real)%.17g}, ", cont ? 1 : 0, has_aff ? 1 : 0,
                 has_neg ? 1 : 0, rootfind, ip, c.save_before ? 1 : 0, c.save_after ? 1 : 0, abstol, nudge);
        table += buf;
        cond += "        case " + std::to_string(k) + ": return " + c.condition_name + "(u, p, B200_USER_T(t));\n";   // reverse-time programs: the caller's t
        aff += "        case " + std::to_string(k) + ": ";
        if (has_neg) aff += std::string("if (neg) { ") + c.affect_neg_name + "(u, p, B200_USER_T(t), terminate); break; } ";
        if (has_aff) aff += std::string("if (!neg) { ") + c.affect_name + "(u, p, B200_USER_T(t), terminate); } ";
        aff += "break;\n";
    }
    table += "}\n";
    out += "#define B200_CALLBACKS 1\n#define B200_NCB " + std::to_string(ncb) + "\n#define B200_NCC " + std::to_string(ncc) + "\n";
    out += table;
    out += "__device__ __forceinline__ real b200_cb_condition(int k, const real* u, const real* p, real t) {\n    switch (k) {\n" + cond +
           "    }\n    return (real)0;\n}\n";
    out += "__device__ __forceinline__ void b200_cb_affect(int k, bool neg, real* u, real* p, real t, int* terminate) {\n    switch (k) {\n" + aff +
           "    }\n}\n";
    return B200ODE_OK;
}

// Assemble the translation unit and run NVRTC.
int nvrtc_build(int alg, int dtype, int n, int np, const char* rhs_src, const char* rhs_name,
                const char* jac_src, const char* jac_name, const char* tgrad_src, const char* tgrad_name,
                const char* extra_options, std::vector<char>& cubin, std::string& log, double* ms, int* coop_l = nullptr,
                const B200CallbackSrc* cbs = nullptr, int ncb = 0, size_t* dyn_smem_out = nullptr, int* wide_nt_out = nullptr) {
    int rc = validate_compile_args(alg, dtype, n, np, rhs_src, rhs_name, jac_src, jac_name, tgrad_src, tgrad_name);
    if (rc) return rc;
    std::string cb_text;
    if (ncb > 0) {
        const bool everystep = extra_options && strstr(extra_options, "-DB200_EVERYSTEP=1");
        rc = callbacks_source(alg, dtype, cbs, ncb, everystep, extra_options && strstr(extra_options, "-DB200_COOP=1"), cb_text);
        if (rc) return rc;
    }
    bool stiff = is_stiff_alg(alg);
    auto t_begin = std::chrono::steady_clock::now();

    std::string tu;
    tu += "#define B200_N " + std::to_string(n) + "\n";
    tu += "#define B200_NP " + std::to_string(np) + "\n";
    tu += "#define B200_F32 " + std::to_string(dtype == B200ODE_F32 ? 1 : 0) + "\n";
    tu += "#define B200_ALG " + std::to_string(alg) + "\n";
    tu += "#include \"b200_base.cuh\"\n";
    // forward declarations make the user's functions __device__ __forceinline__
    // (their definitions carry no execution-space annotation: -default-device).
    // -DB200_RHS_INLINE=0 keeps the RHS out of line (one copy instead of one per stage).  Measured
    // on Pleiades/Vern7 (n = 28): the out-of-line / rolled variants are slower (230-310 ms vs
    // 177-220 ms for 2^18 trajectories) although the fully inlined loop body (470 KB) misses the
    // instruction cache — dynamic indexing forces every stage vector through local memory.
    // Default: inline unless the inlined copies would add up to more source than ptxas digests in reasonable time
    // (Vern9 x Pleiades, 28 call sites x 6.5 KB: > 15 min inlined, 22 s out of line).
    int call_sites = 2;     // b200_initdt
    switch (alg) {
        case B200ODE_ALG_TSIT5: case B200ODE_ALG_DP5: call_sites += 7; break;
        case B200ODE_ALG_BS3: call_sites += 4; break;
        case B200ODE_ALG_VERN6: call_sites += 12; break;
        case B200ODE_ALG_VERN7: call_sites += 16; break;
        case B200ODE_ALG_VERN8: call_sites += 21; break;
        case B200ODE_ALG_VERN9: call_sites += 26; break;
        case B200ODE_ALG_ROSENBROCK23: case B200ODE_ALG_ROSENBROCK32: call_sites += 3; break;
        case B200ODE_ALG_AUTOTSIT5_ROSENBROCK23: call_sites += 12; break;
        case B200ODE_ALG_RODAS5P: case B200ODE_ALG_RODAS5PE: case B200ODE_ALG_RODAS5: call_sites += 8; break;
        default: call_sites += 6; break;
    }
    bool rhs_inline = strlen(rhs_src) * (size_t)call_sites <= 120000;
    if (extra_options && strstr(extra_options, "-DB200_RHS_INLINE=1")) rhs_inline = true;
    if (extra_options && strstr(extra_options, "-DB200_RHS_INLINE=0")) rhs_inline = false;
    // Lane-group kernel (device/b200_coop.cuh): opt-in with -DB200_COOP=1; the RHS is then in component form
    //     real NAME(int i, const real* u, const real* p, const real t)   returning du_i
    const bool coop = extra_options && strstr(extra_options, "-DB200_COOP=1");
    // Shared-memory stage kernel (device/b200_vern7_wide.cuh): opt-in with -DB200_WIDE=1; the RHS keeps the
    // full-vector form and is inlined once, compiled in the two arithmetic flavours described below
    const bool wide = extra_options && strstr(extra_options, "-DB200_WIDE=1");
    if (wide && (coop || alg != B200ODE_ALG_VERN7))
        return fail(B200ODE_EUNSUPPORTED, "the shared-memory stage kernel (B200ODE_OPT_SMEM_STAGES) is available for Vern7 in the one-thread form");
    if (wide && ncb > 0)
        return fail(B200ODE_EUNSUPPORTED, "callbacks are not available in the shared-memory stage kernel");
    // Reverse time (B200ODE_OPT_REVERSE_TIME; tdir = -1, solve.jl:273): the kernels integrate du/ds = -f(u, p, -s) over
    // (-t0, -tf) — see device/b200_ensemble.cuh.  The user's functions are wrapped below.
    const bool reverse = extra_options && strstr(extra_options, "-DB200_REVERSE=1");
    if (reverse && (coop || wide))
        return fail(B200ODE_EUNSUPPORTED, "reverse-time integration is served by the one-thread-per-trajectory kernel");
    if (reverse && extra_options && strstr(extra_options, "-DB200_TSPANS=1"))
        return fail(B200ODE_EUNSUPPORTED, "reverse-time integration is not combined with per-trajectory time spans");
    if (coop && alg != B200ODE_ALG_VERN7 && alg != B200ODE_ALG_ROSENBROCK23)
        return fail(B200ODE_EUNSUPPORTED, "the lane-group kernel (B200ODE_OPT_COMPONENT_RHS) is available for Vern7 and Rosenbrock23");
    if (coop && alg == B200ODE_ALG_ROSENBROCK23 && (n < 2 || n > 16 || n == 3))
        return fail(B200ODE_EUNSUPPORTED, "the lane-group Rosenbrock23 (warp-shuffle LU) serves n = 2 and 4..16; n = 3 uses the in-register "
                                          "inverse of the one-thread kernel");
    if (!coop && !wide)
    tu += std::string(rhs_inline ? "__device__ __forceinline__ void " : "__device__ __noinline__ void ") + rhs_name +
          "(real* du, const real* u, const real* p, const real t);\n";
    if (stiff && !coop) {
        tu += std::string("__device__ __forceinline__ void ") + jac_name +
              "(real* J, const real* u, const real* p, const real t);\n";
        if (tgrad_src)
            tu += std::string("__device__ __forceinline__ void ") + tgrad_name +
                  "(real* dT, const real* u, const real* p, const real t);\n";
    }
    tu += "// ---- user source (RHS) ----\n";
    if (coop) {
        // Component form: the user's text is compiled twice as a member function of two wrapper structs.
        //  B200UserFast:  sqrt()/sqrtf() and the optional B200_DIV(a, b) hint expand to the flagged branch-free
        //                 sequences of b200_base.cuh (bit-identical to the IEEE operators unless they raise the flag),
        //                 so the independent terms of a sum interleave instead of queueing behind one another's
        //                 slow-path branches;
        //  B200UserExact: the plain operators; evaluated only when the fast evaluation raised its flag.
        // (Rosenbrock23: the Jacobian entry function and the optional time-gradient component function ride along)
        std::string body = strip_includes(rhs_src);
        if (stiff) {
            body += strip_includes(jac_src);
            if (tgrad_src) body += strip_includes(tgrad_src);
        }
        tu += "struct B200UserFast {\n  bool b200_bad;\n"
              "#define B200_DIV(a, b) b200_div_fast((a), (b), b200_bad)\n"
              "#define sqrt(x) b200_sqrt_fast((x), b200_bad)\n#define sqrtf(x) b200_sqrt_fast((x), b200_bad)\n";
        tu += body;
        tu += "\n#undef B200_DIV\n#undef sqrt\n#undef sqrtf\n};\n";
        tu += "struct B200UserExact {\n#define B200_DIV(a, b) ((a) / (b))\n";
        tu += body;
        tu += "\n#undef B200_DIV\n};\n";
        tu += std::string("#define B200_USER_COMP_NAME ") + rhs_name + "\n";
        if (stiff) {
            tu += std::string("#define B200_USER_JAC_NAME ") + jac_name + "\n";
            if (tgrad_src) tu += std::string("#define B200_USER_TGRAD_NAME ") + tgrad_name + "\n";
        }
    } else if (wide) {
        // Full-vector form, compiled twice as member functions (see the component form above): B200UserFast with
        // sqrt / B200_DIV mapped to the flagged branch-free sequences, B200UserExact with the plain operators
        const std::string body = strip_includes(rhs_src);
        // (B200_WIDE_WINDOW: see b200_sqrt_window in device/b200_vern7_wide.cuh — the text's roots are started at most
        //  that many root/quotient groups ahead, so the register allocator is not flooded by the scheduler)
        tu += "#ifndef B200_WIDE_WINDOW\n#define B200_WIDE_WINDOW 0\n#endif\n"
              "#include \"b200_window.cuh\"\n"
              "struct B200UserFast {\n  bool b200_bad;\n  bool b200_win[B200_WIDE_WINDOW + 1];\n"
              "#define B200_DIV(a, b) b200_div_fast((a), (b), b200_bad)\n"
              "#define sqrt(x) b200_sqrt_window((x), b200_bad, b200_win)\n#define sqrtf(x) b200_sqrt_window((x), b200_bad, b200_win)\n"
              "__device__ __forceinline__\n";
        tu += body;
        tu += "\n#undef B200_DIV\n#undef sqrt\n#undef sqrtf\n};\n";
        tu += "struct B200UserExact {\n#define B200_DIV(a, b) ((a) / (b))\n__device__ __forceinline__\n";
        tu += body;
        tu += "\n#undef B200_DIV\n};\n";
        tu += std::string("#define B200_USER_FULL_NAME ") + rhs_name + "\n";
    } else
    tu += strip_includes(rhs_src);
    if (stiff && !coop) {
        tu += "// ---- user source (Jacobian) ----\n";
        tu += strip_includes(jac_src);
        if (tgrad_src) {
            tu += "// ---- user source (time gradient) ----\n";
            tu += strip_includes(tgrad_src);
        }
    }
    tu += cb_text;
    tu += "// ---- steppers ----\n";
    if (coop) tu += "#define B200_USER_RHS_COMP(i,u,p,t) (B200UserExact().B200_USER_COMP_NAME((i),(u),(p),(t)))\n";
    else if (wide) tu += "#define B200_USER_RHS(du,u,p,t) (B200UserExact().B200_USER_FULL_NAME((du),(u),(p),(t)))\n";
    else if (reverse) {
        // g(u, p, s) = -f(u, p, -s);  dg/du = -J(u, p, -s);  dg/ds = +f_t(u, p, -s)  (negation is exact)
        tu += std::string("__device__ __forceinline__ void b200_rev_rhs(real* du, const real* u, const real* p, const real s) {\n    ") +
              rhs_name + "(du, u, p, -s);\n#pragma unroll\n    for (int i = 0; i < B200_N; ++i) du[i] = -du[i];\n}\n"
              "#define B200_USER_RHS(du,u,p,t) b200_rev_rhs((du),(u),(p),(t))\n";
        if (stiff) {
            tu += std::string("__device__ __forceinline__ void b200_rev_jac(real* J, const real* u, const real* p, const real s) {\n    ") +
                  jac_name + "(J, u, p, -s);\n#pragma unroll\n    for (int i = 0; i < B200_N * B200_N; ++i) J[i] = -J[i];\n}\n"
                  "#define B200_JAC(J,u,p,t) b200_rev_jac((J),(u),(p),(t))\n";
            if (tgrad_src) tu += std::string("#define B200_TGRAD(dT,u,p,t) ") + tgrad_name + "((dT),(u),(p),-(t))\n";
        }
    }
    else tu += std::string("#define B200_USER_RHS(du,u,p,t) ") + rhs_name + "((du),(u),(p),(t))\n";
    if (stiff && !coop && !reverse) {
        tu += std::string("#define B200_JAC(J,u,p,t) ") + jac_name + "((J),(u),(p),(t))\n";
        if (tgrad_src) tu += std::string("#define B200_TGRAD(dT,u,p,t) ") + tgrad_name + "((dT),(u),(p),(t))\n";
    }
    tu += "#include \"b200_ensemble.cuh\"\n";

    nvrtcProgram prog;
    nvrtcResult r = nvrtcCreateProgram(&prog, tu.c_str(), "b200_ensemble_tu.cu", k_b200_num_headers,
                                       k_b200_header_sources, k_b200_header_names);
    if (r != NVRTC_SUCCESS) return fail(B200ODE_ECOMPILE, std::string("nvrtcCreateProgram: ") + nvrtcGetErrorString(r));

    std::vector<std::string> opts = {
        "--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "-default-device",
        "-lineinfo", "--ptxas-options=-v",
    };
    // launch bounds: small systems keep everything in registers at 4 CTAs x 128 threads per SM
    bool has_block = false, has_minb = false;
    if (extra_options) {
        std::string eo(extra_options);
        size_t pos = 0;
        while (pos < eo.size()) {
            size_t sp = eo.find(' ', pos);
            std::string tok = eo.substr(pos, sp == std::string::npos ? std::string::npos : sp - pos);
            if (!tok.empty()) {
                opts.push_back(tok);
                if (tok.rfind("-DB200_BLOCK=", 0) == 0) has_block = true;
                if (tok.rfind("-DB200_MINBLOCKS=", 0) == 0) has_minb = true;
            }
            if (sp == std::string::npos) break;
            pos = sp + 1;
        }
    }
    int words = n * (dtype == B200ODE_F32 ? 1 : 2);
    int coop_lanes = 0;
    if (coop) {
        // lanes per trajectory: the smallest power of two that leaves at most 2 components per lane (n = 28 -> 16)
        const char* at = extra_options ? strstr(extra_options, "-DB200_L=") : nullptr;
        if (at) coop_lanes = atoi(at + 9);
        else {
            coop_lanes = 2;
            const int per_lane = stiff ? 1 : 2;       // the shuffle LU keeps one row per lane
            while (coop_lanes < 32 && (n + coop_lanes - 1) / coop_lanes > per_lane) coop_lanes *= 2;
            opts.push_back("-DB200_L=" + std::to_string(coop_lanes));
        }
        if (coop_lanes < 2 || coop_lanes > 32 || (coop_lanes & (coop_lanes - 1)))
            return fail(B200ODE_EINVAL, "-DB200_L= must be a power of two in 2..32");
    }
    if (coop_l) *coop_l = coop_lanes;
    if (wide) {
        // threads per CTA that own a trajectory: 9 stage slots x n reals each, inside the 227 KB a CTA may opt in to
        // on sm_100; a multiple of 16 (a half-filled last warp still adds trajectories), at most 512
        const size_t per_thread = (size_t)9 * n * (dtype == B200ODE_F32 ? 4 : 8);
        long long nt = (long long)(232448 / per_thread);
        const char* at = strstr(extra_options, "-DB200_WIDE_NT=");
        if (at) nt = std::min<long long>(nt, atoll(at + 15));
        else nt = std::min<long long>(nt, 512) & ~15ll;
        if (nt < 16) return fail(B200ODE_EUNSUPPORTED, "the shared-memory stage kernel needs at least 16 trajectories per SM: n is too large");
        if (!at) opts.push_back("-DB200_WIDE_NT=" + std::to_string(nt));
        if (!has_block) { opts.push_back("-DB200_BLOCK=" + std::to_string((nt + 31) / 32 * 32)); has_block = true; }
        if (!has_minb) { opts.push_back("-DB200_MINBLOCKS=1"); has_minb = true; }
        // software-pipelining window of the RHS's roots (device/b200_window.cuh): 3..4 measured best for Pleiades
        if (!strstr(extra_options, "-DB200_WIDE_WINDOW=")) opts.push_back("-DB200_WIDE_WINDOW=4");
        if (dyn_smem_out) *dyn_smem_out = per_thread * (size_t)nt;
        if (wide_nt_out) *wide_nt_out = (int)nt;
    }
    // measured launch shapes (scripts/sweep_dev.py): small explicit systems run best as one 512-thread
    // CTA per SM at 128 registers (FP64) / three 256-thread CTAs at 80 registers (FP32)
    const bool small_explicit = !stiff && words <= 8 &&
                                (alg == B200ODE_ALG_TSIT5 || alg == B200ODE_ALG_DP5 || alg == B200ODE_ALG_BS3);
    int real_cbs = 0;
    for (int i = 0; i < ncb; ++i) real_cbs += (cbs[i].kind != B200ODE_CB_ISOUTOFDOMAIN);
    if (real_cbs > 0 && !has_block && !has_minb) {       // event handling needs registers beyond the step loop's
        opts.push_back("-DB200_BLOCK=128"); opts.push_back("-DB200_MINBLOCKS=3");
        has_block = has_minb = true;
    }
    if (small_explicit && !has_block && !has_minb) {
        if (dtype == B200ODE_F32) { opts.push_back("-DB200_BLOCK=256"); opts.push_back("-DB200_MINBLOCKS=3"); }
        else { opts.push_back("-DB200_BLOCK=512"); opts.push_back("-DB200_MINBLOCKS=1"); }
        has_block = has_minb = true;
    }
    // Rosenbrock-type steppers on small FP64 systems: one 512-thread CTA per SM at 128 registers, no spills (measured after
    // the divisions were regrouped, scripts/sweep_rober.py, 2^20 Robertson trajectories: Rodas5P 7.84 ms against 8.03 ms at
    // four 128-thread CTAs; Rosenbrock23 7.68 ms against 8.34 ms at five)
    if (stiff && !coop && alg != B200ODE_ALG_AUTOTSIT5_ROSENBROCK23 && dtype == B200ODE_F64 && words <= 8 && real_cbs == 0 &&
        !has_block && !has_minb) {
        opts.push_back("-DB200_BLOCK=512"); opts.push_back("-DB200_MINBLOCKS=1");
        has_block = has_minb = true;
    }
    if (coop && !has_minb) { opts.push_back("-DB200_MINBLOCKS=4"); has_minb = true; }
    if (!has_block) opts.push_back("-DB200_BLOCK=128");
    if (!has_minb) {
        // registers available per thread at k CTAs of 128 threads: 65536/(128k)
        int minb = 4;
        if (alg == B200ODE_ALG_AUTOTSIT5_ROSENBROCK23) minb = (words <= 8) ? 3 : 1;
        else if (stiff) minb = (words <= 8) ? ((alg == B200ODE_ALG_ROSENBROCK23 || alg == B200ODE_ALG_ROSENBROCK32) ? 5 : 4) : 1;   // measured (scripts/sweep_rober.py)
        else if (alg == B200ODE_ALG_VERN7 || alg == B200ODE_ALG_VERN6 || alg == B200ODE_ALG_VERN8 || alg == B200ODE_ALG_VERN9)
            minb = (words <= 6) ? (alg == B200ODE_ALG_VERN9 ? 2 : 3) : 1;
        else minb = (words <= 8) ? 4 : (words <= 16 ? 2 : 1);
        opts.push_back("-DB200_MINBLOCKS=" + std::to_string(minb));
    }
    // Staged saveat queue (b200_stage_rows in device/b200_ensemble.cuh; B200ODE_OPT_STAGED_SAVEAT): Tsit5 programs with a
    // rectangular output may pack their saveat rows through a per-warp shared-memory queue.  Opt-in: measured slower than
    // the in-step loop on the headline workload.
    {
        const bool off = extra_options && strstr(extra_options, "-DB200_STAGE_ROWS=0");
        const bool on_req = extra_options && strstr(extra_options, "-DB200_STAGE_ROWS=1");
        const bool eligible = alg == B200ODE_ALG_TSIT5 && !coop && !wide && ncb == 0 && n <= 4 &&
                              !(extra_options && (strstr(extra_options, "-DB200_EVERYSTEP=1") || strstr(extra_options, "-DB200_SAVE_IDXS=")));
        if (on_req && !eligible) { nvrtcDestroyProgram(&prog); return fail(B200ODE_EUNSUPPORTED, "-DB200_STAGE_ROWS=1 needs Tsit5, n <= 4, no save_everystep / save_idxs / callbacks"); }
        (void)off;
        if (eligible && on_req) {
            int block = 128, minb = 1;
            for (auto& o : opts) {
                if (o.rfind("-DB200_BLOCK=", 0) == 0) block = atoi(o.c_str() + 13);
                if (o.rfind("-DB200_MINBLOCKS=", 0) == 0) minb = atoi(o.c_str() + 17);
            }
            const size_t rs = dtype == B200ODE_F32 ? 4 : 8;
            const size_t per_warp = ((size_t)(8 * n + 2) * 56 * rs + 64 * (rs + 12) + 16 + 15) / 16 * 16;
            const size_t bytes = per_warp * (size_t)(block / 32);
            if (bytes * (size_t)minb <= 226 * 1024) {
                if (!on_req) opts.push_back("-DB200_STAGE_ROWS=1");
                if (dyn_smem_out) *dyn_smem_out = bytes;
            } else if (on_req) { nvrtcDestroyProgram(&prog); return fail(B200ODE_EUNSUPPORTED, "-DB200_STAGE_ROWS=1: the row queues do not fit the SM's shared memory at this launch shape"); }
        }
    }
    std::vector<const char*> copts;
    for (auto& o : opts) copts.push_back(o.c_str());
    r = nvrtcCompileProgram(prog, (int)copts.size(), copts.data());
    size_t log_size = 0;
    nvrtcGetProgramLogSize(prog, &log_size);
    log.assign(log_size > 0 ? log_size : 1, '\0');
    if (log_size > 0) nvrtcGetProgramLog(prog, &log[0]);
    while (!log.empty() && log.back() == '\0') log.pop_back();
    if (r != NVRTC_SUCCESS) {
        nvrtcDestroyProgram(&prog);
        return fail(B200ODE_ECOMPILE, std::string("NVRTC: ") + nvrtcGetErrorString(r) + "\n" + log);
    }
    size_t sz = 0;
    r = nvrtcGetCUBINSize(prog, &sz);
    if (r != NVRTC_SUCCESS || sz == 0) {
        nvrtcDestroyProgram(&prog);
        return fail(B200ODE_ECOMPILE, "NVRTC produced no cubin");
    }
    cubin.resize(sz);
    nvrtcGetCUBIN(prog, cubin.data());
    nvrtcDestroyProgram(&prog);
    if (ms) *ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return B200ODE_OK;
}

// ---------------------------------------------------------------------------
// AOT kernels (compiled by nvcc into this library)

// deterministic per-component sum: fixed grid, fixed tree order
template <typename R>
__global__ void __launch_bounds__(256) k_reduce_partial(const R* __restrict__ x, long long ts, long long cs,
                                                        long long count, int n, double* __restrict__ partial) {
    __shared__ double sh[256];
    for (int c = 0; c < n; ++c) {
        double acc = 0.0;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
             i += (long long)gridDim.x * blockDim.x)
            acc += (double)x[i * ts + c * cs];
        sh[threadIdx.x] = acc;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) partial[(size_t)blockIdx.x * n + c] = sh[0];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_reduce_final(const double* __restrict__ partial, int nblocks, int n,
                                                      double* __restrict__ out) {
    __shared__ double sh[256];
    for (int c = 0; c < n; ++c) {
        double acc = 0.0;
        for (int b = threadIdx.x; b < nblocks; b += blockDim.x) acc += partial[(size_t)b * n + c];
        sh[threadIdx.x] = acc;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[c] = sh[0];
        __syncthreads();
    }
}

// column statistics of a row-major [rows][cols] matrix: per-CTA partial sums in fixed order
template <typename R>
__global__ void __launch_bounds__(256) k_colsum_partial(const R* __restrict__ x, long long rows, int cols,
                                                        const double* __restrict__ mean, double* __restrict__ partial) {
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const double m = mean ? mean[c] : 0.0;
    double acc = 0.0;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const double v = (double)x[r * (long long)cols + c];
        if (mean) { const double d = v - m; acc = fma(d, d, acc); }
        else acc += v;
    }
    partial[(size_t)blockIdx.x * cols + c] = acc;
}
__global__ void __launch_bounds__(256) k_colsum_final(const double* __restrict__ partial, int nparts, int cols, double scale,
                                                      double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    double acc = 0.0;
    for (int b = 0; b < nparts; ++b) acc += partial[(size_t)b * cols + c];
    out[c] = acc * scale;
}

// FMA-pipe peak: 8 independent dependent chains per thread, register resident
// ---- exclusive scan of the per-trajectory row counts (save_everystep): int32[N] -> int64[N+1] -------------
// Three small launches: per-tile sums (1024 counts per CTA), a single-CTA scan of the tile sums, and the
// per-tile exclusive scan offset by its tile base.  HBM-bound and tiny next to the integration passes; it only
// exists so that the counts never travel to the host (8 bytes do: the total).
#define B200_SCAN_TILE 1024
__global__ void __launch_bounds__(256) k_scan_tile_sums(const int* __restrict__ counts, long long N, long long* __restrict__ tile_sums) {
    __shared__ long long sm[256];
    const long long base = (long long)blockIdx.x * B200_SCAN_TILE;
    long long acc = 0;
    for (int k = 0; k < 4; ++k) {
        long long i = base + threadIdx.x * 4 + k;
        if (i < N) acc += counts[i];
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sm[threadIdx.x] += sm[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = sm[0];
}
__global__ void __launch_bounds__(1024) k_scan_tiles(long long* __restrict__ tile_sums, int ntiles, long long* __restrict__ total) {
    // single CTA, sequential over chunks of 1024 tiles (ntiles is N/1024: a few thousand at most)
    __shared__ long long sm[1024];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int c0 = 0; c0 < ntiles; c0 += 1024) {
        const int i = c0 + (int)threadIdx.x;
        const long long v = (i < ntiles) ? tile_sums[i] : 0;
        sm[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {          // Hillis-Steele inclusive scan
            long long add = ((int)threadIdx.x >= off) ? sm[threadIdx.x - off] : 0;
            __syncthreads();
            sm[threadIdx.x] += add;
            __syncthreads();
        }
        if (i < ntiles) tile_sums[i] = carry + sm[threadIdx.x] - v;    // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry += sm[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(256) k_scan_apply(const int* __restrict__ counts, long long N, const long long* __restrict__ tile_base,
                                                    const long long* __restrict__ total, long long* __restrict__ offsets) {
    __shared__ long long sm[256];
    const long long base = (long long)blockIdx.x * B200_SCAN_TILE;
    long long v[4], acc = 0;
    for (int k = 0; k < 4; ++k) {
        long long i = base + threadIdx.x * 4 + k;
        v[k] = (i < N) ? counts[i] : 0;
        acc += v[k];
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        long long add = ((int)threadIdx.x >= off) ? sm[threadIdx.x - off] : 0;
        __syncthreads();
        sm[threadIdx.x] += add;
        __syncthreads();
    }
    long long run = tile_base[blockIdx.x] + sm[threadIdx.x] - acc;
    for (int k = 0; k < 4; ++k) {
        long long i = base + threadIdx.x * 4 + k;
        if (i < N) offsets[i] = run;
        run += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) offsets[N] = *total;
}

// ---- self-test of the flagged fast division / square root (device/b200_base.cuh) against the plain operators ----
__device__ __forceinline__ unsigned long long st_mix(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// operand generator: class 0 = arbitrary bit patterns (every exponent, subnormals, Inf, NaN), 1 = moderate exponents
// (what an ODE step sees), 2 = mantissas near all-ones / all-zeros (the hard cases of Markstein-type corrections)
__device__ __forceinline__ double st_f64(unsigned long long r, int cls) {
    if (cls == 0) return __longlong_as_double((long long)r);
    unsigned long long mant = r & 0xFFFFFFFFFFFFFull, sign = r & 0x8000000000000000ull;
    unsigned long long e = 1023ull - 40ull + ((r >> 52) % 81ull);
    if (cls == 2) { const unsigned k = (unsigned)((r >> 40) & 31u); mant = ((r >> 47) & 1ull) ? (0xFFFFFFFFFFFFFull >> k << k) : (mant >> (k + 20)); }
    return __longlong_as_double((long long)(sign | (e << 52) | mant));
}
__device__ __forceinline__ float st_f32(unsigned long long r, int cls) {
    const unsigned u = (unsigned)(r >> 17);
    if (cls == 0) return __uint_as_float(u);
    unsigned mant = u & 0x7FFFFFu, sign = u & 0x80000000u, e = 127u - 30u + ((unsigned)(r >> 50) % 61u);
    if (cls == 2) { const unsigned k = (unsigned)((r >> 8) & 15u); mant = ((r >> 7) & 1ull) ? (0x7FFFFFu >> k << k) : (mant >> (k + 6)); }
    return __uint_as_float(sign | (e << 23) | mant);
}
__global__ void __launch_bounds__(256) k_selftest_fastmath(long long samples, unsigned long long seed,
                                                           unsigned long long* counters) {
    // counters: [0] div64 [1] div_const64 [2] sqrt64 [3] div32 [4] sqrt32 [5] unguarded div32 (mismatches); [6] flagged
    unsigned long long bad[6] = {0, 0, 0, 0, 0, 0}, flag_count = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < samples; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long r0 = st_mix(seed + 3ull * (unsigned long long)i), r1 = st_mix(r0), r2 = st_mix(r1);
        const int cls = (int)(r2 % 3ull);
        {   // binary64 division
            const double a = st_f64(r0, cls), b = st_f64(r1, cls);
            bool f = false;
            const double q = b200_div_fast(a, b, f), e = a / b;
            flag_count += f;
            if (!f && __double_as_longlong(q) != __double_as_longlong(e) && !(q != q && e != e)) bad[0]++;
            // a / c with rc = RN(1/c) (b200_div_const_fast) for the divisor classes the kernels use: gamma = 0.9, the
            // state dimension n = 1..64, and binary32-valued divisors (fastpower results).  (For arbitrary divisors
            // the 3-operation form is NOT always the IEEE quotient — Markstein's exceptions — so it is never used there.)
            const unsigned sel = (unsigned)(r2 >> 60) % 3u;
            const double c = sel == 0 ? 0.9 : (sel == 1 ? (double)(1 + (int)((r2 >> 20) & 63ull))
                                                        : (double)(0.001f + 999.0f * (float)((r1 >> 11) & 0xFFFFFFu) * (1.0f / 16777216.0f)));
            bool g = false;
            const double qc = b200_div_const_fast(a, c, 1.0 / c, g);
            if (!g && __double_as_longlong(qc) != __double_as_longlong(a / c)) bad[1]++;
        }
        {   // binary64 square root
            const double x = fabs(st_f64(r2, cls));
            bool f = false;
            const double s = b200_sqrt_fast(x, f), e = sqrt(x);
            flag_count += f;
            if (!f && __double_as_longlong(s) != __double_as_longlong(e) && !(s != s && e != e)) bad[2]++;
        }
        {   // binary32 division and square root
            const float a = st_f32(r0, cls), b = st_f32(r1, cls), x = fabsf(st_f32(r2, cls));
            bool f = false;
            const float q = b200_div_fast(a, b, f), e = a / b;
            if (!f && __float_as_uint(q) != __float_as_uint(e) && !(q != q && e != e)) bad[3]++;
            bool g = false;
            const float s = b200_sqrt_fast(x, g), es = sqrtf(x);
            flag_count += f + g;
            if (!g && __float_as_uint(s) != __float_as_uint(es) && !(s != s && es != es)) bad[4]++;
            // the unguarded division of fastlog2: den in [1.27, 2.03), |num| < 1.3
            const float den = 1.27f + 0.76f * (float)((r1 >> 11) & 0xFFFFFFu) * (1.0f / 16777216.0f);
            const float num = -0.55f + 1.85f * (float)((r0 >> 11) & 0xFFFFFFu) * (1.0f / 16777216.0f);
            if (__float_as_uint(b200_div_fast_nocheck(num, den)) != __float_as_uint(num / den)) bad[5]++;
        }
    }
    for (int k = 0; k < 6; ++k) if (bad[k]) atomicAdd(counters + k, bad[k]);
    if (flag_count) atomicAdd(counters + 6, flag_count);
}

template <typename R>
__global__ void __launch_bounds__(256) k_fma_peak(R* out, int iters, R a, R b) {
    R x0 = (R)threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    R s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == (R)123456789) out[0] = s;   // never true; keeps the chains alive
}

// set further down, next to compile_impl (which lives in the C-linkage part of this file)
int (*g_compile_nosave)(b200ode_program) = nullptr;

template <typename R>
int launch_solve_fwd(b200ode_handle h, b200ode_program prog, const B200DeviceProblem* dp, const B200Opts* o,
                     B200DeviceResult* dr, cudaStream_t stream, const long long* row_offsets, void* ts_rag, void* dts_rag);

// Reverse-time programs (prog->reverse, tf < t0): the launch is that of the mirrored problem — span (-t0, -tf), every time
// list negated (the descending saveat list becomes ascending), |dt| — the kernels hand times back through B200_USER_T.
template <typename R>
int launch_solve(b200ode_handle h, b200ode_program prog, const B200DeviceProblem* dp, const B200Opts* o,
                 B200DeviceResult* dr, cudaStream_t stream, const long long* row_offsets = nullptr, void* ts_rag = nullptr,
                 void* dts_rag = nullptr) {
    if (!prog->reverse) return launch_solve_fwd<R>(h, prog, dp, o, dr, stream, row_offsets, ts_rag, dts_rag);
    B200DeviceProblem dm = *dp;
    dm.t0 = -dp->t0; dm.tf = -dp->tf;
    B200Opts om = *o;
    auto mirrored = [](const double* v, int nv) {
        std::vector<double> out(v && nv > 0 ? nv : 0);
        for (size_t i = 0; i < out.size(); ++i) out[i] = -v[i];
        return out;
    };
    const std::vector<double> sa = mirrored(o->saveat, o->nsaveat), st = mirrored(o->tstops, o->ntstops),
                              dd = mirrored(o->d_discontinuities, o->nd_discontinuities);
    if (!sa.empty()) om.saveat = sa.data();
    if (!st.empty()) om.tstops = st.data();
    if (!dd.empty()) om.d_discontinuities = dd.data();
    om.dt = std::fabs(o->dt);         // a positive dt is converted, a negative one is the direction's own (solve.jl:981-983)
    om.dtmax = std::fabs(o->dtmax);   // likewise dtmax (solve.jl:401: a positive dtmax is converted); 0 stays the default
    return launch_solve_fwd<R>(h, prog, &dm, &om, dr, stream, row_offsets, ts_rag, dts_rag);
}

template <typename R>
int launch_solve_fwd(b200ode_handle h, b200ode_program prog, const B200DeviceProblem* dp, const B200Opts* o,
                     B200DeviceResult* dr, cudaStream_t stream, const long long* row_offsets, void* ts_rag, void* dts_rag) {
    const int n = prog->n, np = prog->np;
    const long long N = dp->trajectories;
    Params<R> P{};
    P.N = N;
    P.u0 = (const R*)dp->u0;
    if (dp->u0_shared) { P.u0_ts = 0; P.u0_cs = 1; }
    else if (dp->u0_layout == B200ODE_LAYOUT_SOA) { P.u0_ts = 1; P.u0_cs = N; }
    else { P.u0_ts = n; P.u0_cs = 1; }
    P.p = (const R*)dp->p;
    if (dp->p_shared) { P.p_ts = 0; P.p_cs = 1; }
    else if (dp->p_layout == B200ODE_LAYOUT_SOA) { P.p_ts = 1; P.p_cs = N; }
    else { P.p_ts = np; P.p_cs = 1; }
    P.t0 = (R)dp->t0; P.tf = (R)dp->tf;
    P.tspans = dp->tspans; P.dtmax_default = (o->dtmax > 0) ? 0 : 1;
    if ((dp->tspans != nullptr) != prog->tspans)
        return fail(B200ODE_EINVAL, prog->tspans ? "a program compiled with -DB200_TSPANS=1 needs problem.tspans"
                                                 : "problem.tspans needs a program compiled with -DB200_TSPANS=1");
    if (prog->tspans && dr->us && !prog->everystep)
        return fail(B200ODE_EUNSUPPORTED, "per-trajectory time spans: the rectangular `us` output is not available (use the ragged output)");
    P.reltol = (R)(o->reltol > 0 ? o->reltol : 1e-3);
    P.abstol = (R)(o->abstol > 0 ? o->abstol : 1e-6);
    if ((o->abstol_vec || o->reltol_vec) && !prog->vector_tol)
        return fail(B200ODE_EINVAL, "opts.abstol_vec / reltol_vec need a program compiled with B200ODE_OPT_VECTOR_TOL");
    if (prog->vector_tol) {
        // reltol[0..n) then abstol[0..n) in the real type; a missing vector is the scalar broadcast.  Module-level constant
        // memory: re-uploaded (after a device synchronisation) only when the values change.
        std::vector<double> key(2 * (size_t)n);
        for (int i = 0; i < n; ++i) {
            key[i] = o->reltol_vec ? o->reltol_vec[i] : (o->reltol > 0 ? o->reltol : 1e-3);
            key[n + i] = o->abstol_vec ? o->abstol_vec[i] : (o->abstol > 0 ? o->abstol : 1e-6);
        }
        if (key != prog->tolv_cached) {
            std::vector<R> vals(key.begin(), key.end());
            void* sym = nullptr; size_t bytes = 0;
            CUDA_TRY(cudaLibraryGetGlobal(&sym, &bytes, prog->lib, "B200_TOLV"));
            if (bytes != sizeof(R) * vals.size()) return fail(B200ODE_ECUDA, "B200_TOLV has an unexpected size");
            CUDA_TRY(cudaDeviceSynchronize());
            CUDA_TRY(cudaMemcpy(sym, vals.data(), bytes, cudaMemcpyHostToDevice));
            prog->tolv_cached = key;
        }
    }
    P.dt_user = (R)o->dt;
    P.dtmin = (R)o->dtmin;
    P.dtmax = (R)(o->dtmax > 0 ? o->dtmax : (dp->tf - dp->t0));
    P.maxiters = o->maxiters > 0 ? o->maxiters : 1000000;
    P.nsaveat = o->saveat ? o->nsaveat : 0;
    P.save_start = (o->save_start != 0) ? 1 : 0;
    P.save_end = (o->save_end < 0) ? 1 : (o->save_end ? 2 : 0);
    B200Problem hp{}; hp.t0 = dp->t0; hp.tf = dp->tf;
    P.nslots = (dr->us && !prog->everystep) ? nslots_typed(&hp, o, prog->dtype) : 0;
    P.row_offsets = row_offsets; P.ts_rag = (R*)ts_rag; P.dts_rag = (R*)dts_rag;
    P.tstops = nullptr; P.ntstops = 0;
    const bool want_tstops = o->tstops && o->ntstops > 0;
    const bool want_disc = o->d_discontinuities && o->nd_discontinuities > 0;
    P.disc = nullptr; P.ndisc = 0;
    if (want_tstops && !prog->tstops) return fail(B200ODE_EINVAL, "opts.tstops needs a program compiled with -DB200_TSTOPS=1");
    if (want_disc && !prog->tstops) return fail(B200ODE_EINVAL, "opts.d_discontinuities needs a program compiled with -DB200_TSTOPS=1");
    if (want_disc && prog->callbacks) return fail(B200ODE_EUNSUPPORTED, "d_discontinuities are not combined with callbacks");
    if (prog->tstops) {
        // initialize_tstops: stops strictly inside (t0, tf), ascending, duplicates kept, tf last — in the real type
        std::vector<R> stops;
        for (int i = 0; want_tstops && i < o->ntstops; ++i) {
            const R v = (R)o->tstops[i];
            if (v > (R)dp->t0 && v < (R)dp->tf) stops.push_back(v);
        }
        // ... and the d_discontinuities inside (t0, tf) (initialize_tstops, solve.jl:1033-1036)
        std::vector<R> discs;
        for (int i = 0; want_disc && i < o->nd_discontinuities; ++i) {
            const R v = (R)o->d_discontinuities[i];
            if (v > (R)dp->t0 && v < (R)dp->tf) stops.push_back(v);
            if (v >= (R)dp->t0) discs.push_back(v);          // reinit_d_discontinuities! (solve.jl:1185-1197)
        }
        std::sort(discs.begin(), discs.end());
        std::sort(stops.begin(), stops.end());
        stops.push_back((R)dp->tf);
        const size_t nstops = stops.size();
        stops.insert(stops.end(), discs.begin(), discs.end());      // one device buffer: [stops..., discontinuities...]
        std::vector<double> key(stops.begin(), stops.end());
        key.push_back((double)nstops);
        if (h->tstops_cached_dtype != (int)sizeof(R) || key != h->tstops_cached) {      // re-uploaded only when it changes
            CUDA_TRY(cudaDeviceSynchronize());      // no launch may still be reading the old list
            CUDA_TRY(h->tstops.ensure(sizeof(R) * stops.size()));
            CUDA_TRY(cudaMemcpy(h->tstops.ptr, stops.data(), sizeof(R) * stops.size(), cudaMemcpyHostToDevice));
            h->tstops_cached = key; h->tstops_cached_dtype = (int)sizeof(R);
        }
        P.tstops = (const R*)h->tstops.ptr; P.ntstops = (int)nstops;
        P.disc = P.tstops + nstops; P.ndisc = (int)discs.size();
    }
    P.saveat = nullptr;
    if (P.nsaveat > 0) {
        // the grid travels as real[] in the handle's scratch; re-uploaded only when it changes
        const bool same = (h->saveat_cached_dtype == (int)sizeof(R)) && (int)h->saveat_cached.size() == P.nsaveat &&
                          std::equal(h->saveat_cached.begin(), h->saveat_cached.end(), o->saveat);
        if (!same) {
            std::vector<R> grid(P.nsaveat);
            for (int i = 0; i < P.nsaveat; ++i) grid[i] = (R)o->saveat[i];
            CUDA_TRY(cudaDeviceSynchronize());          // no launch may still be reading the old grid
            CUDA_TRY(h->saveat.ensure(sizeof(R) * P.nsaveat));
            CUDA_TRY(cudaMemcpy(h->saveat.ptr, grid.data(), sizeof(R) * P.nsaveat, cudaMemcpyHostToDevice));
            h->saveat_cached.assign(o->saveat, o->saveat + P.nsaveat);
            h->saveat_cached_dtype = (int)sizeof(R);
        }
        P.saveat = (const R*)h->saveat.ptr;
    }
    // Initial step sizes: b200_initdt writes them into the caller's t_final[] (each trajectory reads its own slot when it
    // starts and overwrites it with t_final when it ends), so the launch owns no handle-level scratch and launches on
    // different streams cannot disturb each other.
    P.dt0 = (R*)dr->t_final;
    P.npeer = 0; P.peer_world = 1; P.peer_rank = 0; P.peer_block = 1;
    for (int r = 0; r < 8; ++r) P.peer_out[r] = nullptr;
    if (dr->npeers > 0) {
        if (dr->npeers > 8 || dr->peer_world < 1 || dr->peer_rank < 0 || dr->peer_rank >= dr->peer_world || dr->peer_block < 1)
            return fail(B200ODE_EINVAL, "result.peer_*: npeers in 1..8, 0 <= peer_rank < peer_world, peer_block >= 1");
        if (prog->coop_l > 0 || prog->wide_nt > 0)
            return fail(B200ODE_EUNSUPPORTED, "the fused peer gather is served by the one-thread kernels");
        for (int r = 0; r < dr->npeers; ++r) {
            if (!dr->peer_u_final[r]) return fail(B200ODE_EINVAL, "result.peer_u_final holds a NULL pointer");
            P.peer_out[r] = (R*)dr->peer_u_final[r];
        }
        P.npeer = dr->npeers; P.peer_world = dr->peer_world; P.peer_rank = dr->peer_rank; P.peer_block = dr->peer_block;
    }
    P.u_final = (R*)dr->u_final;
    if (dr->u_final_layout == B200ODE_LAYOUT_SOA) { P.uf_ts = 1; P.uf_cs = N; } else { P.uf_ts = n; P.uf_cs = 1; }
    P.t_final = (R*)dr->t_final;
    P.us = (R*)dr->us;
    P.naccept = dr->naccept; P.nreject = dr->nreject; P.nf = dr->nf; P.retcode = dr->retcode; P.nsaved = dr->nsaved;
    P.njacs = dr->njacs; P.nw = dr->nw; P.nsolve = dr->nsolve;
    // work counter: one of a ring of kCounterRing counters per launch (zeroed on the launch's own stream), so up to
    // kCounterRing launches may be in flight on one handle
    CUDA_TRY(h->counter.ensure(sizeof(unsigned long long) * kCounterRing));
    P.work_counter = (unsigned long long*)h->counter.ptr + (h->launch_seq.fetch_add(1) % kCounterRing);
    P.flags = o->flags;
    {   // tstop tolerance 100*eps(max(|t|,|tf|)) (integrator_utils.jl:277-286) is constant when |t0| <= |tf|
        const R at0 = std::fabs(P.t0), atf = std::fabs(P.tf);
        P.tol_const = (!(at0 > atf) && dp->tspans == nullptr) ? 1 : 0;
        P.tol100_tf = (R)100 * (std::nextafter(atf, std::numeric_limits<R>::infinity()) - atf);
    }
    {   // controller start state: fastpower(qoldinit, beta2), beta2 = 2//(5 order) (4//100 for DP5) rounded to the real type
        const R beta2 = (R)(prog->alg == B200ODE_ALG_DP5 ? 4.0 / 100.0 : 2.0 / (5.0 * alg_order(prog->alg)));
        P.fpe0 = b200_fastpower((R)1e-4, beta2);
        P.rfpe0 = (R)1 / P.fpe0;
    }
    CUDA_TRY(cudaMemsetAsync(P.work_counter, 0, sizeof(unsigned long long), stream));

    if (!prog->adaptive && o->dt == 0.0 && !want_tstops)
        return fail(B200ODE_EINVAL, "Fixed timestep methods require a choice of dt or choosing the tstops");   // solve.jl:277-280
    void* args[] = {&P};
    if (o->dt == 0.0 && prog->adaptive) {
        unsigned g = (unsigned)((N + 255) / 256);
        CUDA_TRY(cudaLaunchKernel((const void*)prog->k_initdt, dim3(g), dim3(256), args, 0, stream));
    }
    unsigned grid, small_block = 0;
    if (prog->wide_nt > 0) {
        if (P.nsaveat > 0) {
            // interior rows need the interpolant, whose lazy stages k11..k16 this variant does not store
            for (int i = 0; i < o->nsaveat; ++i)
                if ((R)o->saveat[i] > (R)dp->t0 && (R)o->saveat[i] < (R)dp->tf)
                    return fail(B200ODE_EUNSUPPORTED, "saveat points inside (t0, tf) are not available in the shared-memory stage kernel "
                                                      "(B200ODE_OPT_SMEM_STAGES): compile the program without it");
        }
        const long long need = (N + prog->wide_nt - 1) / prog->wide_nt;
        grid = (unsigned)std::min<long long>(prog->info.grid, need > 0 ? need : 1);
    } else if (prog->coop_l > 0) {
        const long long per_cta = (long long)(prog->info.block / prog->coop_l);      // trajectories in flight per CTA
        const long long need = (N + per_cta - 1) / per_cta;
        grid = (unsigned)std::min<long long>(prog->info.grid, need > 0 ? need : 1);
    } else if (o->flags & B200ODE_FLAG_STATIC_SCHEDULE) grid = (unsigned)((N + prog->info.block - 1) / prog->info.block);
    else {
        grid = (unsigned)prog->info.grid;
        long long need = (N + prog->info.block - 1) / prog->info.block;
        if ((long long)grid > need) grid = (unsigned)(need > 0 ? need : 1);
        // Small ensembles (BASELINE configs[0]: 10 k trajectories): full CTAs would occupy N / block SMs with four warps per
        // scheduler each while the other SMs idle; the launch bound is only a maximum, so the CTAs shrink until the
        // trajectories spread over all SMs (one or two warps per scheduler finish a step in its latency, not in 4x it)
        if (need < (long long)h->num_sms) {
            const long long per_sm = (N + h->num_sms - 1) / h->num_sms;
            small_block = (unsigned)std::max<long long>(32, (per_sm + 31) / 32 * 32);
            if (small_block < (unsigned)prog->info.block) grid = (unsigned)std::max<long long>(1, (N + small_block - 1) / small_block);
            else small_block = 0;
        }
    }
    b200ode_program run = prog;
    if (P.nsaveat == 0 && prog->nosave_ok) {       // the build without saveat handling (made on first use)
        std::lock_guard<std::mutex> lk(prog->nosave_mu);
        if (!prog->nosave_tried) { prog->nosave_tried = true; if (g_compile_nosave) g_compile_nosave(prog); }
        if (prog->nosave) run = prog->nosave;
    }
    CUDA_TRY(cudaLaunchKernel((const void*)run->k_integrate, dim3(grid), dim3(small_block ? small_block : run->info.block), args,
                              run->dyn_smem, stream));
    return B200ODE_OK;
}

int check_problem(int64_t N, const void* u0, const void* p, int np, double t0, double tf, const B200Opts* o, bool reverse = false) {
    if (N < 0) return fail(B200ODE_EINVAL, "trajectories must be >= 0");
    if (!u0) return fail(B200ODE_EINVAL, "u0 is NULL");
    if (np > 0 && !p) return fail(B200ODE_EINVAL, "p is NULL but the program has np > 0");
    if (!reverse && tf < t0)
        return fail(B200ODE_EUNSUPPORTED, "tf < t0 needs a program compiled with B200ODE_OPT_REVERSE_TIME");
    if (reverse && tf > t0)
        return fail(B200ODE_EINVAL, "a program compiled with B200ODE_OPT_REVERSE_TIME integrates reversed spans (tf < t0)");
    if (!(tf > t0) && !(tf < t0)) return fail(B200ODE_EINVAL, "tspan must have tf != t0 (finite)");
    // (the reference integrates towards an infinite tf until the solution blows up, test/InterfaceI/inf_handling.jl; the kernels'
    //  stop tolerance 100 eps(max(|t|, |tf|)) is written for finite spans, so an infinite span is declined, not mis-stepped)
    if (!std::isfinite(t0) || !std::isfinite(tf)) return fail(B200ODE_EUNSUPPORTED, "tspan must be finite");
    if (!o) return fail(B200ODE_EINVAL, "opts is NULL");
    if (o->nsaveat < 0) return fail(B200ODE_EINVAL, "nsaveat < 0");
    if (o->saveat) {
        const double td = reverse ? -1.0 : 1.0;          // the list is given in the order the integrator meets it
        double prev = td * t0;
        for (int i = 0; i < o->nsaveat; ++i) {
            double s = td * o->saveat[i];
            if (!(s >= prev) || !(s > td * t0) || !(s <= td * tf))
                return fail(B200ODE_EINVAL, reverse ? "saveat must be descending with every entry in [tf, t0) (reverse time)"
                                                    : "saveat must be ascending with every entry in (t0, tf]");
            prev = s;
        }
    }
    if (o->ntstops < 0 || (o->ntstops > 0 && !o->tstops)) return fail(B200ODE_EINVAL, "bad tstops");
    for (int i = 0; i < o->ntstops; ++i)
        if (!std::isfinite(o->tstops[i])) return fail(B200ODE_EINVAL, "tstops must be finite");
    if (o->dt < 0 && !reverse) return fail(B200ODE_EINVAL, "dt must be >= 0 (0 = automatic)");
    if (o->dtmin < 0) return fail(B200ODE_EINVAL, "dtmin must be >= 0");
    return B200ODE_OK;
}

}  // namespace

// ===========================================================================
extern "C" {

const char* b200ode_version(void) { return "b200ode 0.1.0 (sm_100a, NVRTC " "12.x" ")"; }
int b200ode_struct_size(int which) {
    switch (which) {
        case 0: return (int)sizeof(B200Problem);
        case 1: return (int)sizeof(B200Opts);
        case 2: return (int)sizeof(B200Result);
        case 3: return (int)sizeof(B200DeviceProblem);
        case 4: return (int)sizeof(B200DeviceResult);
        case 5: return (int)sizeof(B200ProgramInfo);
        case 6: return (int)sizeof(B200CallbackSrc);
        case 7: return (int)sizeof(B200Ragged);
        default: return -1;
    }
}

const char* b200ode_last_error(b200ode_handle) { return g_last_error.c_str(); }

int b200ode_create(b200ode_handle* out, int device_id) {
    if (!out) return fail(B200ODE_EINVAL, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(B200ODE_ECUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                       "); this library has no CPU fallback");
    if (device_id < 0 || device_id >= count) return fail(B200ODE_EINVAL, "device_id out of range");
    CUDA_TRY(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device_id));
    if (prop.major != 10)
        return fail(B200ODE_EUNSUPPORTED, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                              "; this library contains sm_100a code only");
    b200ode_handle h = new b200ode_handle_s();
    h->device = device_id;
    h->num_sms = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&h->ev0)); CUDA_TRY(cudaEventCreate(&h->ev1));
    CUDA_TRY(cudaEventCreate(&h->ev2)); CUDA_TRY(cudaEventCreate(&h->ev3));
    *out = h;
    return B200ODE_OK;
}

int b200ode_destroy(b200ode_handle h) {
    if (!h) return B200ODE_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (DevBuf* b : {&h->counter, &h->dt0, &h->saveat, &h->scratch_t, &h->in_u0, &h->in_p, &h->out_uf, &h->out_tf,
                      &h->out_us, &h->out_i32, &h->red_partial, &h->stat_out, &h->row_offsets, &h->rag_dts, &h->dense_tq, &h->dense_out, &h->scan_tiles, &h->tstops, &h->in_tspans})
        b->release();
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (cudaEvent_t e : {h->ev0, h->ev1, h->ev2, h->ev3}) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->chunk_events) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) { if (h->stage[i]) cudaFreeHost(h->stage[i]); if (h->stage_ev[i]) cudaEventDestroy(h->stage_ev[i]); }
    delete h;
    return B200ODE_OK;
}

void b200ode_free(void* p) { free(p); }

int b200ode_compile_only(int alg, int dtype, int n, int np, const char* rhs_src, const char* rhs_name,
                         const char* jac_src, const char* jac_name, const char* tgrad_src, const char* tgrad_name,
                         const char* extra_options, void** cubin, size_t* cubin_bytes, char** log) {
    std::vector<char> cb; std::string lg;
    if (cubin) *cubin = nullptr;
    if (cubin_bytes) *cubin_bytes = 0;
    if (log) *log = nullptr;
    int rc = nvrtc_build(alg, dtype, n, np, rhs_src, rhs_name, jac_src, jac_name, tgrad_src, tgrad_name,
                         extra_options, cb, lg, nullptr);
    if (log) {
        const std::string& src = (rc == B200ODE_ECOMPILE) ? g_last_error : lg;
        *log = (char*)malloc(src.size() + 1);
        memcpy(*log, src.c_str(), src.size() + 1);
    }
    if (rc) return rc;
    if (cubin) { *cubin = malloc(cb.size()); memcpy(*cubin, cb.data(), cb.size()); }
    if (cubin_bytes) *cubin_bytes = cb.size();
    return B200ODE_OK;
}

static int compile_impl(b200ode_handle h, b200ode_program* out, int alg, int dtype, int n, int np,
                        const char* rhs_src, const char* rhs_name, const char* jac_src, const char* jac_name,
                        const char* tgrad_src, const char* tgrad_name, const char* extra_options,
                        const B200CallbackSrc* cbs, int ncb);
int b200ode_compile(b200ode_handle h, b200ode_program* out, int alg, int dtype, int n, int np,
                    const char* rhs_src, const char* rhs_name, const char* jac_src, const char* jac_name,
                    const char* tgrad_src, const char* tgrad_name, const char* extra_options) {
    return compile_impl(h, out, alg, dtype, n, np, rhs_src, rhs_name, jac_src, jac_name, tgrad_src, tgrad_name, extra_options,
                        nullptr, 0);
}
int b200ode_compile_callbacks(b200ode_handle h, b200ode_program* out, int alg, int dtype, int n, int np,
                              const char* rhs_src, const char* rhs_name, const char* jac_src, const char* jac_name,
                              const char* tgrad_src, const char* tgrad_name, const B200CallbackSrc* callbacks, int ncallbacks,
                              const char* extra_options) {
    if (ncallbacks < 0 || (ncallbacks > 0 && !callbacks)) return fail(B200ODE_EINVAL, "callbacks is NULL");
    return compile_impl(h, out, alg, dtype, n, np, rhs_src, rhs_name, jac_src, jac_name, tgrad_src, tgrad_name, extra_options,
                        callbacks, ncallbacks);
}
int b200ode_compile_only_callbacks(int alg, int dtype, int n, int np, const char* rhs_src, const char* rhs_name,
                                   const B200CallbackSrc* callbacks, int ncallbacks, const char* extra_options,
                                   void** cubin, size_t* cubin_bytes, char** log) {
    std::vector<char> cb; std::string lg;
    if (cubin) *cubin = nullptr;
    if (cubin_bytes) *cubin_bytes = 0;
    if (log) *log = nullptr;
    int rc = nvrtc_build(alg, dtype, n, np, rhs_src, rhs_name, nullptr, nullptr, nullptr, nullptr, extra_options, cb, lg, nullptr,
                         nullptr, callbacks, ncallbacks);
    if (log) {
        const std::string& src = (rc != B200ODE_OK) ? g_last_error : lg;
        *log = (char*)malloc(src.size() + 1);
        memcpy(*log, src.c_str(), src.size() + 1);
    }
    if (rc) return rc;
    if (cubin) { *cubin = malloc(cb.size()); memcpy(*cubin, cb.data(), cb.size()); }
    if (cubin_bytes) *cubin_bytes = cb.size();
    return B200ODE_OK;
}
static int compile_impl(b200ode_handle h, b200ode_program* out, int alg, int dtype, int n, int np,
                        const char* rhs_src, const char* rhs_name, const char* jac_src, const char* jac_name,
                        const char* tgrad_src, const char* tgrad_name, const char* extra_options,
                        const B200CallbackSrc* cbs, int ncb) {
    if (!h || !out) return fail(B200ODE_EINVAL, "handle/out is NULL");
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(h->device));
    b200ode_program prog = new b200ode_program_s();
    prog->h = h; prog->alg = alg; prog->dtype = dtype; prog->n = n; prog->np = np;
    std::string log; double ms = 0;
    int rc = nvrtc_build(alg, dtype, n, np, rhs_src, rhs_name, jac_src, jac_name, tgrad_src, tgrad_name,
                         extra_options, prog->cubin, log, &ms, &prog->coop_l, cbs, ncb, &prog->dyn_smem, &prog->wide_nt);
    if (rc) { delete prog; return rc; }
    prog->callbacks = false;
    for (int i = 0; i < ncb; ++i) prog->callbacks = prog->callbacks || (cbs[i].kind != B200ODE_CB_ISOUTOFDOMAIN);
    prog->vector_tol = extra_options && strstr(extra_options, "-DB200_VECTOR_TOL=1");
    if (prog->vector_tol && prog->coop_l > 0) { delete prog; return fail(B200ODE_EUNSUPPORTED, "per-component tolerances are not available in the lane-group kernel"); }
    prog->everystep = extra_options && strstr(extra_options, "-DB200_EVERYSTEP=1");
    prog->tstops = extra_options && strstr(extra_options, "-DB200_TSTOPS=1");
    prog->adaptive = !(extra_options && strstr(extra_options, "-DB200_ADAPTIVE=0"));
    prog->tspans = extra_options && strstr(extra_options, "-DB200_TSPANS=1");
    prog->reverse = extra_options && strstr(extra_options, "-DB200_REVERSE=1");
    if (prog->tspans && (prog->tstops || prog->coop_l > 0 || prog->wide_nt > 0 || ncb > 0)) {
        delete prog; return fail(B200ODE_EUNSUPPORTED, "per-trajectory time spans are not combined with tstops / d_discontinuities, callbacks, "
                                                       "the lane-group or the shared-memory stage kernel");
    }
    if ((!prog->adaptive || prog->tstops || prog->everystep) && prog->coop_l > 0) {
        delete prog; return fail(B200ODE_EUNSUPPORTED, "adaptive=false, tstops and save_everystep are not available in the lane-group kernel");
    }
    prog->nsave = parse_save_idxs(extra_options, n);
    if (prog->nsave < 0) { delete prog; return fail(B200ODE_EINVAL, "-DB200_SAVE_IDXS= must list 0-based component indices below n, comma separated"); }
    if (prog->nsave == 0) prog->nsave = n;
    if (prog->everystep && prog->wide_nt > 0) { delete prog; return fail(B200ODE_EUNSUPPORTED, "save_everystep is not available in the shared-memory stage kernel"); }
    if (prog->nsave != n && prog->coop_l > 0) { delete prog; return fail(B200ODE_EUNSUPPORTED, "save_idxs is not available in the lane-group kernel"); }
    cudaError_t e = cudaLibraryLoadData(&prog->lib, prog->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) { delete prog; return fail(B200ODE_ECUDA, std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e)); }
    e = cudaLibraryGetKernel(&prog->k_integrate, prog->lib, "b200_integrate");
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&prog->k_initdt, prog->lib, "b200_initdt");
    if (e == cudaSuccess && prog->everystep && prog->nsave == n && alg != B200ODE_ALG_ROSENBROCK32 &&
        alg != B200ODE_ALG_AUTOTSIT5_ROSENBROCK23 && !prog->callbacks)
        e = cudaLibraryGetKernel(&prog->k_dense, prog->lib, "b200_dense_eval");
    if (e != cudaSuccess) {
        cudaLibraryUnload(prog->lib); delete prog;
        return fail(B200ODE_ECUDA, std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(e));
    }
    cudaFuncAttributes fa{};
    e = cudaFuncGetAttributes(&fa, (const void*)prog->k_integrate);
    if (e != cudaSuccess) { cudaLibraryUnload(prog->lib); delete prog; return fail(B200ODE_ECUDA, std::string("cudaFuncGetAttributes: ") + cudaGetErrorString(e)); }
    prog->info.regs_integrate = fa.numRegs;
    prog->info.local_bytes_integrate = (int)fa.localSizeBytes;
    prog->info.smem_bytes_integrate = (int)fa.sharedSizeBytes;
    prog->info.block = fa.maxThreadsPerBlock > 0 ? std::min(fa.maxThreadsPerBlock, 1024) : 128;
    // maxThreadsPerBlock reflects __launch_bounds__(B200_BLOCK, …)
    cudaFuncAttributes fb{};
    if (cudaFuncGetAttributes(&fb, (const void*)prog->k_initdt) == cudaSuccess) {
        prog->info.regs_initdt = fb.numRegs;
        prog->info.local_bytes_initdt = (int)fb.localSizeBytes;
    }
    if (prog->dyn_smem > 48 * 1024) {
        e = cudaFuncSetAttribute((const void*)prog->k_integrate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prog->dyn_smem);
        if (e != cudaSuccess) { cudaLibraryUnload(prog->lib); delete prog; return fail(B200ODE_ECUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)); }
    }
    prog->info.smem_bytes_integrate += (int)prog->dyn_smem;
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)prog->k_integrate, prog->info.block, prog->dyn_smem);
    if (e != cudaSuccess || nb < 1) nb = 1;
    prog->info.blocks_per_sm = nb;
    prog->info.grid = nb * h->num_sms;
    prog->info.cubin_bytes = (int64_t)prog->cubin.size();
    prog->info.compile_ms = ms;
    {
        const bool has = extra_options && strstr(extra_options, "-DB200_NO_SAVEAT=");
        const bool staged = extra_options && strstr(extra_options, "-DB200_STAGE_ROWS=1");
        prog->nosave_ok = !has && !staged && ncb == 0 && prog->coop_l == 0 && prog->wide_nt == 0 && !prog->everystep && !prog->vector_tol;
        if (prog->nosave_ok) {
            prog->keep_rhs = rhs_src; prog->keep_rhs_name = rhs_name;
            if (jac_src) { prog->keep_jac = jac_src; prog->keep_jac_name = jac_name; }
            if (tgrad_src) { prog->keep_tg = tgrad_src; prog->keep_tg_name = tgrad_name; }
            prog->keep_extra = extra_options ? extra_options : "";
        }
    }
    *out = prog;
    return B200ODE_OK;
}

// the same program with -DB200_NO_SAVEAT=1 (same launch shape: the shape options are derived from alg / dtype / n)
static int compile_nosave_variant(b200ode_program prog) {
    b200ode_program alt = nullptr;
    const std::string extra = prog->keep_extra.empty() ? std::string("-DB200_NO_SAVEAT=1") : prog->keep_extra + " -DB200_NO_SAVEAT=1";
    const int rc = compile_impl(prog->h, &alt, prog->alg, prog->dtype, prog->n, prog->np, prog->keep_rhs.c_str(), prog->keep_rhs_name.c_str(),
                                prog->keep_jac.empty() ? nullptr : prog->keep_jac.c_str(), prog->keep_jac.empty() ? nullptr : prog->keep_jac_name.c_str(),
                                prog->keep_tg.empty() ? nullptr : prog->keep_tg.c_str(), prog->keep_tg.empty() ? nullptr : prog->keep_tg_name.c_str(),
                                extra.c_str(), nullptr, 0);
    if (rc == B200ODE_OK && alt && alt->info.block == prog->info.block) prog->nosave = alt;
    else if (alt) b200ode_program_destroy(alt);
    return rc;
}
static const bool g_nosave_registered = (g_compile_nosave = &compile_nosave_variant, true);

int b200ode_program_destroy(b200ode_program prog) {
    if (!prog) return B200ODE_OK;
    if (prog->nosave) { b200ode_program_destroy(prog->nosave); prog->nosave = nullptr; }
    if (prog->lib) cudaLibraryUnload(prog->lib);
    delete prog;
    return B200ODE_OK;
}

int b200ode_program_info(b200ode_program prog, B200ProgramInfo* info) {
    if (!prog || !info) return fail(B200ODE_EINVAL, "NULL argument");
    *info = prog->info;
    return B200ODE_OK;
}

// rows per trajectory as the KERNEL counts them: grid values and tf compared in the program's real type (a grid point
// that rounds to (float)tf is the end point for an F32 program), and with save_end = false EVERY grid entry equal to tf
// is skipped (skip_saveat_at_tspan_end), not just one
}  // extern "C"
namespace { int nslots_typed(const B200Problem* prob, const B200Opts* o, int dtype) {
    if (!prob || !o || !o->saveat || o->nsaveat <= 0) return 0;
    const int save_start = (o->save_start != 0) ? 1 : 0;
    const int save_end = (o->save_end != 0) ? 1 : 0;
    auto is_tf = [&](double v) { return dtype == B200ODE_F32 ? ((float)v == (float)prob->tf) : (v == prob->tf); };
    int at_tf = 0;
    for (int i = 0; i < o->nsaveat; ++i) at_tf += is_tf(o->saveat[i]) ? 1 : 0;
    int slots = save_start + o->nsaveat;
    if (at_tf > 0 && !save_end) slots -= at_tf;        // skip_saveat_at_tspan_end
    if (at_tf == 0 && save_end) slots += 1;            // solution_endpoint_match_cur_integrator!
    return slots;
} }  // namespace
extern "C" {
int b200ode_nslots(const B200Problem* prob, const B200Opts* o) { return nslots_typed(prob, o, B200ODE_F64); }
int b200ode_nslots_program(b200ode_program prog, const B200Problem* prob, const B200Opts* o) {
    return nslots_typed(prob, o, prog ? prog->dtype : B200ODE_F64);
}

int b200ode_solve_device(b200ode_handle h, b200ode_program prog, const B200DeviceProblem* dp, const B200Opts* o,
                         B200DeviceResult* dr, void* stream) {
    if (!h || !prog || !dp || !o || !dr) return fail(B200ODE_EINVAL, "NULL argument");
    if (prog->h != h) return fail(B200ODE_EINVAL, "program was compiled for a different handle");
    int rc = check_problem(dp->trajectories, dp->u0, dp->p, prog->np, dp->t0, dp->tf, o, prog->reverse);
    if (rc) return rc;
    if (!dr->u_final || !dr->t_final || !dr->naccept || !dr->nreject || !dr->nf || !dr->retcode || !dr->nsaved)
        return fail(B200ODE_EINVAL, "device result arrays u_final,t_final,naccept,nreject,nf,retcode,nsaved are required");
    bool stiff = is_stiff_alg(prog->alg);
    if (stiff && (!dr->njacs || !dr->nw || !dr->nsolve))
        return fail(B200ODE_EINVAL, "Rosenbrock programs need njacs,nw,nsolve result arrays");
    if (prog->everystep) return fail(B200ODE_EINVAL, "program was compiled for save_everystep: use b200ode_solve_everystep[_device]");
    if (dp->trajectories == 0) return B200ODE_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;   // NULL = the legacy default stream
    if (prog->dtype == B200ODE_F32) return launch_solve<float>(h, prog, dp, o, dr, s);
    return launch_solve<double>(h, prog, dp, o, dr, s);
}

int b200ode_solve_everystep_device(b200ode_handle h, b200ode_program prog, const B200DeviceProblem* dp, const B200Opts* o,
                                   B200DeviceResult* dr, const int64_t* row_offsets, void* ts, void* dts, void* stream) {
    if (!h || !prog || !dp || !o || !dr) return fail(B200ODE_EINVAL, "NULL argument");
    if (prog->h != h) return fail(B200ODE_EINVAL, "program was compiled for a different handle");
    if (!prog->everystep) return fail(B200ODE_EINVAL, "program was not compiled with -DB200_EVERYSTEP=1");
    int rc = check_problem(dp->trajectories, dp->u0, dp->p, prog->np, dp->t0, dp->tf, o, prog->reverse);
    if (rc) return rc;
    if (!dr->u_final || !dr->t_final || !dr->naccept || !dr->nreject || !dr->nf || !dr->retcode || !dr->nsaved)
        return fail(B200ODE_EINVAL, "device result arrays u_final,t_final,naccept,nreject,nf,retcode,nsaved are required");
    bool stiff = is_stiff_alg(prog->alg);
    if (stiff && (!dr->njacs || !dr->nw || !dr->nsolve))
        return fail(B200ODE_EINVAL, "Rosenbrock programs need njacs,nw,nsolve result arrays");
    if (row_offsets && (!dr->us || !ts || !dts)) return fail(B200ODE_EINVAL, "the fill pass needs result.us, ts and dts");
    if (dp->trajectories == 0) return B200ODE_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    static_assert(sizeof(long long) == sizeof(int64_t), "row offsets are 64-bit");
    if (prog->dtype == B200ODE_F32) return launch_solve<float>(h, prog, dp, o, dr, s, (const long long*)row_offsets, ts, dts);
    return launch_solve<double>(h, prog, dp, o, dr, s, (const long long*)row_offsets, ts, dts);
}

// Large device->host copy on stream `s`.  Pinned / registered destinations get one asynchronous copy.
// Pageable destinations (plain malloc / numpy memory) would make the driver bounce through its own small
// staging buffer at 2-5 GB/s; instead the copy is pipelined through two pinned 32 MiB buffers owned by the
// handle, with the pinned->pageable memcpy of chunk i (4 host threads) overlapping the DMA of chunk i+1.
// Blocking in the pageable case.
static const size_t kStageBytes = 32u << 20;
static void parallel_memcpy(char* dst, const char* src, size_t bytes) {
    const int T = 4;
    if (bytes < (4u << 20)) { memcpy(dst, src, bytes); return; }
    std::thread th[T];
    size_t per = (bytes + T - 1) / T;
    for (int i = 0; i < T; ++i) {
        size_t o = std::min(bytes, per * i), len = std::min(per, bytes - o);
        th[i] = std::thread([=] { if (len) memcpy(dst + o, src + o, len); });
    }
    for (int i = 0; i < T; ++i) th[i].join();
}
static int d2h_large(b200ode_handle h, void* dst, const void* src, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return B200ODE_OK;
    cudaPointerAttributes attr{};
    bool pinned = (cudaPointerGetAttributes(&attr, dst) == cudaSuccess) && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();     // an unregistered pointer may leave a sticky-less error on old drivers
    if (pinned || bytes < (16u << 20)) {
        CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
        return B200ODE_OK;
    }
    for (int i = 0; i < 2; ++i) {
        if (!h->stage[i]) CUDA_TRY(cudaHostAlloc(&h->stage[i], kStageBytes, cudaHostAllocDefault));
        if (!h->stage_ev[i]) CUDA_TRY(cudaEventCreateWithFlags(&h->stage_ev[i], cudaEventDisableTiming));
    }
    const size_t nchunks = (bytes + kStageBytes - 1) / kStageBytes;
    for (size_t c = 0; c <= nchunks; ++c) {
        if (c < nchunks) {
            const size_t o = c * kStageBytes, len = std::min(kStageBytes, bytes - o);
            CUDA_TRY(cudaMemcpyAsync(h->stage[c & 1], (const char*)src + o, len, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaEventRecord(h->stage_ev[c & 1], s));
        }
        if (c >= 1) {
            const size_t o = (c - 1) * kStageBytes, len = std::min(kStageBytes, bytes - o);
            CUDA_TRY(cudaEventSynchronize(h->stage_ev[(c - 1) & 1]));
            parallel_memcpy((char*)dst + o, (const char*)h->stage[(c - 1) & 1], len);
        }
    }
    return B200ODE_OK;
}

// H2D, counting pass, host exclusive scan, fill pass.  Leaves the ragged rows in the handle's device
// buffers (out_us, scratch_t = ts, rag_dts, row_offsets) and the scalars in out_uf/out_tf/out_i32.
// B200Problem.tspans: validate the pairs, give check_problem the union of the spans, upload the pairs
static int host_tspans(b200ode_handle h, const B200Problem* hp, double* tmin, double* tmax, const double** dev, cudaStream_t s) {
    *dev = nullptr; *tmin = hp->t0; *tmax = hp->tf;
    if (!hp->tspans || hp->trajectories <= 0) return B200ODE_OK;
    double lo = hp->tspans[0], hi = hp->tspans[1];
    for (int64_t i = 0; i < hp->trajectories; ++i) {
        const double a = hp->tspans[2 * i], b = hp->tspans[2 * i + 1];
        if (!std::isfinite(a) || !std::isfinite(b) || !(b > a))
            return fail(B200ODE_EUNSUPPORTED, "tspans: every (t0_i, tf_i) must be finite with tf_i > t0_i (forward time)");
        lo = std::min(lo, a); hi = std::max(hi, b);
    }
    *tmin = lo; *tmax = hi;
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(h->in_tspans.ensure(sizeof(double) * 2 * (size_t)hp->trajectories));
    CUDA_TRY(cudaMemcpyAsync(h->in_tspans.ptr, hp->tspans, sizeof(double) * 2 * (size_t)hp->trajectories, cudaMemcpyHostToDevice, s));
    *dev = (const double*)h->in_tspans.ptr;
    return B200ODE_OK;
}

static int everystep_run(b200ode_handle h, b200ode_program prog, const B200Problem* hp, const B200Opts* o,
                         long long& total_out) {
    const long long N = hp->trajectories;
    const int n = prog->n, np = prog->np;
    const size_t rs = prog->dtype == B200ODE_F32 ? 4 : 8;
    cudaStream_t s = h->stream;
    CUDA_TRY(cudaEventRecord(h->ev0, s));
    size_t u0_bytes = rs * n * (hp->u0_shared ? 1 : (size_t)N);
    size_t p_bytes = rs * np * (hp->p_shared ? 1 : (size_t)N);
    CUDA_TRY(h->in_u0.ensure(u0_bytes));
    CUDA_TRY(cudaMemcpyAsync(h->in_u0.ptr, hp->u0, u0_bytes, cudaMemcpyHostToDevice, s));
    if (np > 0) {
        CUDA_TRY(h->in_p.ensure(p_bytes));
        CUDA_TRY(cudaMemcpyAsync(h->in_p.ptr, hp->p, p_bytes, cudaMemcpyHostToDevice, s));
    }
    CUDA_TRY(h->out_uf.ensure(rs * n * (size_t)N));
    CUDA_TRY(h->out_tf.ensure(rs * (size_t)N));
    CUDA_TRY(h->out_i32.ensure(sizeof(int32_t) * 8 * (size_t)N));
    CUDA_TRY(h->row_offsets.ensure(sizeof(int64_t) * ((size_t)N + 1)));
    int32_t* i32 = (int32_t*)h->out_i32.ptr;
    bool stiff = is_stiff_alg(prog->alg);
    if (!stiff) CUDA_TRY(cudaMemsetAsync(i32 + 4 * N, 0, sizeof(int32_t) * 3 * (size_t)N, s));
    B200DeviceProblem dp{};
    dp.trajectories = N;
    dp.u0 = h->in_u0.ptr; dp.u0_shared = hp->u0_shared; dp.u0_layout = B200ODE_LAYOUT_AOS;
    dp.p = np > 0 ? h->in_p.ptr : nullptr; dp.p_shared = hp->p_shared; dp.p_layout = B200ODE_LAYOUT_AOS;
    dp.t0 = hp->t0; dp.tf = hp->tf;
    {
        double tmin, tmax; const double* dev_tspans = nullptr;
        int rc_t = host_tspans(h, hp, &tmin, &tmax, &dev_tspans, s);
        if (rc_t) return rc_t;
        dp.t0 = tmin; dp.tf = tmax; dp.tspans = dev_tspans;
    }
    B200DeviceResult dr{};
    dr.u_final = h->out_uf.ptr; dr.u_final_layout = B200ODE_LAYOUT_AOS;
    dr.t_final = (double*)h->out_tf.ptr;
    dr.nsaved = i32 + 0 * N; dr.naccept = i32 + 1 * N; dr.nreject = i32 + 2 * N; dr.nf = i32 + 3 * N;
    dr.njacs = i32 + 4 * N; dr.nw = i32 + 5 * N; dr.nsolve = i32 + 6 * N; dr.retcode = i32 + 7 * N;
    CUDA_TRY(cudaEventRecord(h->ev1, s));
    // pass 1: count the rows of every trajectory (the integration is deterministic, so pass 2 repeats it exactly)
    int rc = b200ode_solve_everystep_device(h, prog, &dp, o, &dr, nullptr, nullptr, nullptr, s);
    if (rc) return rc;
    // exclusive scan on the device; only the total comes back (the host needs it to size the buffers)
    const int ntiles = (int)((N + B200_SCAN_TILE - 1) / B200_SCAN_TILE);
    CUDA_TRY(h->scan_tiles.ensure(sizeof(long long) * ((size_t)ntiles + 1)));
    long long* tiles = (long long*)h->scan_tiles.ptr;
    long long* dtotal = tiles + ntiles;
    k_scan_tile_sums<<<ntiles, 256, 0, s>>>(i32, N, tiles);
    k_scan_tiles<<<1, 1024, 0, s>>>(tiles, ntiles, dtotal);
    k_scan_apply<<<ntiles, 256, 0, s>>>(i32, N, tiles, dtotal, (long long*)h->row_offsets.ptr);
    long long total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, dtotal, sizeof(long long), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    total_out = total;
    const size_t rows = (size_t)std::max<long long>(total, 1);
    CUDA_TRY(h->out_us.ensure(rs * (size_t)prog->nsave * rows));
    CUDA_TRY(h->scratch_t.ensure(rs * rows));
    CUDA_TRY(h->rag_dts.ensure(rs * rows));
    // pass 2: fill
    dr.us = h->out_us.ptr;
    rc = b200ode_solve_everystep_device(h, prog, &dp, o, &dr, (const int64_t*)h->row_offsets.ptr, h->scratch_t.ptr,
                                        h->rag_dts.ptr, s);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->ev2, s));
    return B200ODE_OK;
}

// D2H of the per-trajectory scalars + timing, shared by the ragged entry points (stream must be idle afterwards)
static int everystep_finish(b200ode_handle h, b200ode_program prog, long long N, B200Result* res) {
    const int n = prog->n;
    const size_t rs = prog->dtype == B200ODE_F32 ? 4 : 8;
    cudaStream_t s = h->stream;
    int32_t* i32 = (int32_t*)h->out_i32.ptr;
    CUDA_TRY(cudaMemcpyAsync(res->u_final, h->out_uf.ptr, rs * n * (size_t)N, cudaMemcpyDeviceToHost, s));
    std::vector<char> tf_host;
    if (res->t_final) {
        tf_host.resize(rs * (size_t)N);
        CUDA_TRY(cudaMemcpyAsync(tf_host.data(), h->out_tf.ptr, rs * (size_t)N, cudaMemcpyDeviceToHost, s));
    }
    struct { int32_t* dst; int slot; } outs[] = {
        {res->nsaved, 0}, {res->naccept, 1}, {res->nreject, 2}, {res->nf, 3},
        {res->njacs, 4}, {res->nw, 5}, {res->nsolve, 6}, {res->retcode, 7}};
    for (auto& oo : outs)
        if (oo.dst) CUDA_TRY(cudaMemcpyAsync(oo.dst, i32 + (size_t)oo.slot * N, sizeof(int32_t) * (size_t)N, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaEventRecord(h->ev3, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return fail(B200ODE_ECUDA, std::string("kernel failure: ") + cudaGetErrorString(le));
    if (res->t_final) {
        if (rs == 8) memcpy(res->t_final, tf_host.data(), 8 * (size_t)N);
        else for (long long i = 0; i < N; ++i) res->t_final[i] = (double)((const float*)tf_host.data())[i];
    }
    float kms = 0, tms = 0;
    cudaEventElapsedTime(&kms, h->ev1, h->ev2);
    cudaEventElapsedTime(&tms, h->ev0, h->ev3);
    res->kernel_ms = kms; res->total_ms = tms;
    return B200ODE_OK;
}

static int everystep_check(b200ode_handle h, b200ode_program prog, const B200Problem* hp, const B200Opts* o, B200Result* res) {
    if (!h || !prog || !hp || !o || !res) return fail(B200ODE_EINVAL, "NULL argument");
    if (prog->h != h) return fail(B200ODE_EINVAL, "program was compiled for a different handle");
    if (!prog->everystep) return fail(B200ODE_EINVAL, "program was not compiled with -DB200_EVERYSTEP=1");
    double tmin = hp->t0, tmax = hp->tf;
    if (hp->tspans) {
        for (int64_t i = 0; i < hp->trajectories; ++i) {
            const double a = hp->tspans[2 * i], b = hp->tspans[2 * i + 1];
            if (!std::isfinite(a) || !std::isfinite(b) || !(b > a))
                return fail(B200ODE_EUNSUPPORTED, "tspans: every (t0_i, tf_i) must be finite with tf_i > t0_i (forward time)");
            tmin = i == 0 ? a : std::min(tmin, a); tmax = i == 0 ? b : std::max(tmax, b);
        }
    }
    int rc = check_problem(hp->trajectories, hp->u0, hp->p, prog->np, tmin, tmax, o, prog->reverse);
    if (rc) return rc;
    if (!res->u_final) return fail(B200ODE_EINVAL, "result.u_final is required");
    return B200ODE_OK;
}

int b200ode_solve_everystep(b200ode_handle h, b200ode_program prog, const B200Problem* hp, const B200Opts* o, B200Result* res,
                            B200Ragged* out) {
    if (!out) return fail(B200ODE_EINVAL, "NULL argument");
    int rc = everystep_check(h, prog, hp, o, res);
    if (rc) return rc;
    out->total_rows = 0; out->row_offsets = nullptr; out->ts = nullptr; out->us = nullptr;
    const long long N = hp->trajectories;
    if (N == 0) return B200ODE_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    const int n = prog->nsave;      // row width
    const size_t rs = prog->dtype == B200ODE_F32 ? 4 : 8;
    cudaStream_t s = h->stream;
    long long total_ll = 0;
    rc = everystep_run(h, prog, hp, o, total_ll);
    if (rc) return rc;
    const int64_t total = (int64_t)total_ll;
    int64_t* offs_host = (int64_t*)malloc(sizeof(int64_t) * ((size_t)N + 1));
    double* ts_host = (double*)malloc(sizeof(double) * (size_t)std::max<int64_t>(total, 1));
    void* us_host = malloc(rs * (size_t)n * (size_t)std::max<int64_t>(total, 1));
    auto bail = [&](int code) { free(offs_host); free(ts_host); free(us_host); return code; };
    if (!offs_host || !ts_host || !us_host) return bail(fail(B200ODE_EINVAL, "out of host memory"));
    {
        cudaError_t e0 = cudaMemcpyAsync(offs_host, h->row_offsets.ptr, sizeof(int64_t) * ((size_t)N + 1), cudaMemcpyDeviceToHost, s);
        if (e0 != cudaSuccess) return bail(fail(B200ODE_ECUDA, std::string("row_offsets D2H: ") + cudaGetErrorString(e0)));
    }
    std::vector<char> ts_raw;
    rc = d2h_large(h, us_host, h->out_us.ptr, rs * (size_t)n * (size_t)total, s);
    if (rc) return bail(rc);
    if (rs == 8) rc = d2h_large(h, ts_host, h->scratch_t.ptr, 8 * (size_t)total, s);
    else { ts_raw.resize(4 * (size_t)std::max<int64_t>(total, 1)); rc = d2h_large(h, ts_raw.data(), h->scratch_t.ptr, 4 * (size_t)total, s); }
    if (rc) return bail(rc);
    rc = everystep_finish(h, prog, N, res);
    if (rc) return bail(rc);
    if (rs == 4) for (int64_t i = 0; i < total; ++i) ts_host[i] = (double)((const float*)ts_raw.data())[i];
    out->total_rows = total; out->row_offsets = offs_host; out->ts = ts_host; out->us = us_host;
    return B200ODE_OK;
}

extern "C++" {
template <typename R>
struct DenseParams {
    long long N;
    const R* p; long long p_ts, p_cs;
    const long long* row_offsets; const R* ts; const R* dts; const R* us;
    const R* tq; int M;
    R* out;
    R reltol, abstol;
};

template <typename R>
static int launch_dense(b200ode_handle h, b200ode_program prog, long long N, const void* p, int p_shared, int p_layout,
                        const int64_t* row_offsets, const void* ts, const void* dts, const void* us, const void* tq, int M,
                        void* out, const B200Opts* o, cudaStream_t s) {
    DenseParams<R> D{};
    D.N = N;
    D.p = (const R*)p;
    if (p_shared) { D.p_ts = 0; D.p_cs = 1; }
    else if (p_layout == B200ODE_LAYOUT_SOA) { D.p_ts = 1; D.p_cs = N; }
    else { D.p_ts = prog->np; D.p_cs = 1; }
    D.row_offsets = (const long long*)row_offsets; D.ts = (const R*)ts; D.dts = (const R*)dts; D.us = (const R*)us;
    D.tq = (const R*)tq; D.M = M; D.out = (R*)out;
    D.reltol = (R)(o && o->reltol > 0 ? o->reltol : 1e-3);
    D.abstol = (R)(o && o->abstol > 0 ? o->abstol : 1e-6);
    void* args[] = {&D};
    unsigned g = (unsigned)((N + 127) / 128);
    CUDA_TRY(cudaLaunchKernel((const void*)prog->k_dense, dim3(g), dim3(128), args, 0, s));
    return B200ODE_OK;
}
}  // extern "C++"

int b200ode_dense_eval_device(b200ode_handle h, b200ode_program prog, int64_t trajectories, const void* p, int p_shared,
                              int p_layout, const int64_t* row_offsets, const void* ts, const void* dts, const void* us,
                              const void* tq, int nq, void* out, const B200Opts* opts, void* stream) {
    if (!h || !prog || !row_offsets || !ts || !dts || !us || !tq || !out) return fail(B200ODE_EINVAL, "NULL argument");
    if (prog->h != h) return fail(B200ODE_EINVAL, "program was compiled for a different handle");
    if (!prog->everystep || !prog->k_dense) return fail(B200ODE_EINVAL, "program was not compiled with -DB200_EVERYSTEP=1");
    if (prog->np > 0 && !p) return fail(B200ODE_EINVAL, "p is NULL but the program has np > 0");
    if (trajectories <= 0 || nq <= 0) return B200ODE_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (prog->dtype == B200ODE_F32)
        return launch_dense<float>(h, prog, trajectories, p, p_shared, p_layout, row_offsets, ts, dts, us, tq, nq, out, opts, s);
    return launch_dense<double>(h, prog, trajectories, p, p_shared, p_layout, row_offsets, ts, dts, us, tq, nq, out, opts, s);
}

int b200ode_solve_dense(b200ode_handle h, b200ode_program prog, const B200Problem* hp, const B200Opts* o, const double* tq,
                        int nq, void* out, B200Result* res) {
    if (!tq || !out || nq <= 0) return fail(B200ODE_EINVAL, "tq/out is NULL or nq <= 0");
    int rc = everystep_check(h, prog, hp, o, res);
    if (rc) return rc;
    if (o->saveat && o->nsaveat > 0) return fail(B200ODE_EINVAL, "dense output excludes saveat (dense = save_everystep && isempty(saveat), solve.jl:139)");
    if (o->save_start == 0) return fail(B200ODE_EINVAL, "dense output needs save_start");
    if (hp->tspans) return fail(B200ODE_EUNSUPPORTED, "dense output is not available with per-trajectory time spans");
    if (!prog->k_dense) return fail(B200ODE_EUNSUPPORTED, "dense output is not available with save_idxs, for Rosenbrock32 (its stages are not recomputable from the saved rows), for the composite algorithm or with callbacks");
    // the queries come in the order the integration meets them (ode_interpolation sorts by tdir * t, generic_dense.jl:838)
    for (int j = 1; j < nq; ++j)
        if (prog->reverse ? !(tq[j] <= tq[j - 1]) : !(tq[j] >= tq[j - 1]))
            return fail(B200ODE_EINVAL, prog->reverse ? "tq must be descending (reverse time)" : "tq must be ascending");
    const long long N = hp->trajectories;
    if (N == 0) return B200ODE_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    const int n = prog->n;
    const size_t rs = prog->dtype == B200ODE_F32 ? 4 : 8;
    cudaStream_t s = h->stream;
    long long total_ll = 0;
    rc = everystep_run(h, prog, hp, o, total_ll);
    if (rc) return rc;
    CUDA_TRY(h->dense_tq.ensure(rs * (size_t)nq));
    CUDA_TRY(h->dense_out.ensure(rs * (size_t)n * (size_t)nq * (size_t)N));
    std::vector<char> tq_real(rs * (size_t)nq);
    for (int j = 0; j < nq; ++j) { if (rs == 8) ((double*)tq_real.data())[j] = tq[j]; else ((float*)tq_real.data())[j] = (float)tq[j]; }
    CUDA_TRY(cudaMemcpyAsync(h->dense_tq.ptr, tq_real.data(), rs * (size_t)nq, cudaMemcpyHostToDevice, s));
    rc = b200ode_dense_eval_device(h, prog, N, prog->np > 0 ? h->in_p.ptr : nullptr, hp->p_shared, B200ODE_LAYOUT_AOS,
                                   (const int64_t*)h->row_offsets.ptr, h->scratch_t.ptr, h->rag_dts.ptr, h->out_us.ptr,
                                   h->dense_tq.ptr, nq, h->dense_out.ptr, o, s);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->ev2, s));
    rc = d2h_large(h, out, h->dense_out.ptr, rs * (size_t)n * (size_t)nq * (size_t)N, s);
    if (rc) return rc;
    return everystep_finish(h, prog, N, res);
}

static int solve_host_impl(b200ode_handle h, b200ode_program prog, const B200Problem* hp, const B200Opts* o, B200Result* res,
                           double* mean, double* var, bool stats_only);

int b200ode_solve(b200ode_handle h, b200ode_program prog, const B200Problem* hp, const B200Opts* o, B200Result* res) {
    return solve_host_impl(h, prog, hp, o, res, nullptr, nullptr, false);
}

int b200ode_solve_meanvar(b200ode_handle h, b200ode_program prog, const B200Problem* hp, const B200Opts* o, B200Result* res,
                          double* mean, double* var) {
    if (!mean) return fail(B200ODE_EINVAL, "mean is NULL");
    if (!o || !o->saveat || o->nsaveat <= 0) return fail(B200ODE_EINVAL, "timeseries statistics need a saveat grid");
    return solve_host_impl(h, prog, hp, o, res, mean, var, true);
}

static int solve_host_impl(b200ode_handle h, b200ode_program prog, const B200Problem* hp, const B200Opts* o, B200Result* res,
                           double* mean, double* var, bool stats_only) {
    if (!h || !prog || !hp || !o || !res) return fail(B200ODE_EINVAL, "NULL argument");
    if (prog->h != h) return fail(B200ODE_EINVAL, "program was compiled for a different handle");
    double tmin, tmax; const double* dev_tspans = nullptr;
    int rc = host_tspans(h, hp, &tmin, &tmax, &dev_tspans, h->stream);
    if (rc) return rc;
    rc = check_problem(hp->trajectories, hp->u0, hp->p, prog->np, tmin, tmax, o, prog->reverse);
    if (rc) return rc;
    if (hp->tspans && (res->us || stats_only))
        return fail(B200ODE_EUNSUPPORTED, "per-trajectory time spans: the rectangular `us` output and its statistics are not available");
    if (!res->u_final) return fail(B200ODE_EINVAL, "result.u_final is required");
    const long long N = hp->trajectories;
    if (N == 0) return B200ODE_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    const int n = prog->n, np = prog->np;
    const size_t rs = prog->dtype == B200ODE_F32 ? 4 : 8;
    const int nslots = (res->us || stats_only) ? nslots_typed(hp, o, prog->dtype) : 0;
    cudaStream_t s = h->stream;

    CUDA_TRY(cudaEventRecord(h->ev0, s));
    size_t u0_bytes = rs * n * (hp->u0_shared ? 1 : (size_t)N);
    size_t p_bytes = rs * np * (hp->p_shared ? 1 : (size_t)N);
    CUDA_TRY(h->in_u0.ensure(u0_bytes));
    CUDA_TRY(cudaMemcpyAsync(h->in_u0.ptr, hp->u0, u0_bytes, cudaMemcpyHostToDevice, s));
    if (np > 0) {
        CUDA_TRY(h->in_p.ensure(p_bytes));
        CUDA_TRY(cudaMemcpyAsync(h->in_p.ptr, hp->p, p_bytes, cudaMemcpyHostToDevice, s));
    }
    CUDA_TRY(h->out_uf.ensure(rs * n * (size_t)N));
    CUDA_TRY(h->out_tf.ensure(rs * (size_t)N));
    CUDA_TRY(h->out_i32.ensure(sizeof(int32_t) * 8 * (size_t)N));
    const int nsave = prog->nsave;
    size_t us_bytes = rs * (size_t)nsave * (size_t)nslots * (size_t)N;
    if (nslots > 0) CUDA_TRY(h->out_us.ensure(us_bytes));

    int32_t* i32 = (int32_t*)h->out_i32.ptr;
    bool stiff = is_stiff_alg(prog->alg);
    if (!stiff) CUDA_TRY(cudaMemsetAsync(i32 + 4 * N, 0, sizeof(int32_t) * 3 * (size_t)N, s));
    // (rows a failed trajectory does not reach are zero-filled by the kernel itself)

    // Chunked pipeline: the D2H of the saveat rows of chunk c (copy stream) overlaps the
    // kernels of chunk c+1 (compute stream).  Final-state-only solves are a single chunk.
    long long chunk = N;
    if (nslots > 0 && N > 131072 && !stats_only) {
        chunk = (N + 7) / 8;
        chunk = ((chunk + 1023) / 1024) * 1024;
        if (chunk < 65536) chunk = 65536;
    }
    CUDA_TRY(cudaEventRecord(h->ev1, s));
    const size_t row_bytes = rs * (size_t)nsave * (size_t)nslots;
    for (long long c0 = 0; c0 < N; c0 += chunk) {
        const long long cn = std::min(chunk, N - c0);
        B200DeviceProblem dp{};
        dp.trajectories = cn;
        dp.u0 = (char*)h->in_u0.ptr + (hp->u0_shared ? 0 : rs * n * (size_t)c0);
        dp.u0_shared = hp->u0_shared; dp.u0_layout = B200ODE_LAYOUT_AOS;
        dp.p = np > 0 ? (char*)h->in_p.ptr + (hp->p_shared ? 0 : rs * np * (size_t)c0) : nullptr;
        dp.p_shared = hp->p_shared; dp.p_layout = B200ODE_LAYOUT_AOS;
        dp.t0 = tmin; dp.tf = tmax;
        dp.tspans = dev_tspans ? dev_tspans + 2 * (size_t)c0 : nullptr;
        B200DeviceResult dr{};
        dr.u_final = (char*)h->out_uf.ptr + rs * n * (size_t)c0; dr.u_final_layout = B200ODE_LAYOUT_AOS;
        dr.t_final = (double*)((char*)h->out_tf.ptr + rs * (size_t)c0);
        dr.us = nslots > 0 ? (char*)h->out_us.ptr + row_bytes * (size_t)c0 : nullptr;
        dr.nsaved = i32 + 0 * N + c0; dr.naccept = i32 + 1 * N + c0; dr.nreject = i32 + 2 * N + c0; dr.nf = i32 + 3 * N + c0;
        dr.njacs = i32 + 4 * N + c0; dr.nw = i32 + 5 * N + c0; dr.nsolve = i32 + 6 * N + c0; dr.retcode = i32 + 7 * N + c0;
        rc = b200ode_solve_device(h, prog, &dp, o, &dr, s);
        if (rc) return rc;
        if (nslots > 0 && !stats_only) {
            // per-chunk events are created on demand and kept in the handle
            size_t ci = (size_t)(c0 / chunk);
            while (h->chunk_events.size() <= ci) {
                cudaEvent_t e; CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                h->chunk_events.push_back(e);
            }
            CUDA_TRY(cudaEventRecord(h->chunk_events[ci], s));
        }
    }
    if (!stats_only) CUDA_TRY(cudaEventRecord(h->ev2, s));     // end of the last kernel
    if (nslots > 0 && !stats_only) {
        // All chunks are queued; now drain them on the copy stream, each copy gated on its chunk's event.
        // Pinned destinations: every copy is asynchronous and overlaps the kernels of later chunks.
        // Pageable destinations: d2h_large pipelines through the handle's pinned bounce buffers (blocking),
        // still overlapping the kernels that are already in flight.
        for (long long c0 = 0; c0 < N; c0 += chunk) {
            const long long cn = std::min(chunk, N - c0);
            CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->chunk_events[(size_t)(c0 / chunk)], 0));
            rc = d2h_large(h, (char*)res->us + row_bytes * (size_t)c0, (char*)h->out_us.ptr + row_bytes * (size_t)c0,
                           row_bytes * (size_t)cn, h->copy_stream);
            if (rc) return rc;
        }
    }
    if (stats_only) {
        // statistics on the device; only 2 * nslots * n doubles go back
        CUDA_TRY(h->stat_out.ensure(sizeof(double) * 2 * (size_t)nslots * nsave));
        double* dmean = (double*)h->stat_out.ptr;
        double* dvar = dmean + (size_t)nslots * nsave;
        rc = b200ode_timeseries_meanvar_device(h, prog->dtype, h->out_us.ptr, N, nslots, nsave, dmean, var ? dvar : nullptr, s);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(mean, dmean, sizeof(double) * (size_t)nslots * nsave, cudaMemcpyDeviceToHost, s));
        if (var) CUDA_TRY(cudaMemcpyAsync(var, dvar, sizeof(double) * (size_t)nslots * nsave, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaEventRecord(h->ev2, s));
    }

    CUDA_TRY(cudaMemcpyAsync(res->u_final, h->out_uf.ptr, rs * n * (size_t)N, cudaMemcpyDeviceToHost, s));
    std::vector<char> tf_host;
    if (res->t_final) {
        tf_host.resize(rs * (size_t)N);
        CUDA_TRY(cudaMemcpyAsync(tf_host.data(), h->out_tf.ptr, rs * (size_t)N, cudaMemcpyDeviceToHost, s));
    }
    struct { int32_t* dst; int slot; } outs[] = {
        {res->nsaved, 0}, {res->naccept, 1}, {res->nreject, 2}, {res->nf, 3},
        {res->njacs, 4}, {res->nw, 5}, {res->nsolve, 6}, {res->retcode, 7}};
    for (auto& oo : outs)
        if (oo.dst) CUDA_TRY(cudaMemcpyAsync(oo.dst, i32 + (size_t)oo.slot * N, sizeof(int32_t) * (size_t)N, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(h->copy_stream));
    CUDA_TRY(cudaEventRecord(h->ev3, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return fail(B200ODE_ECUDA, std::string("kernel failure: ") + cudaGetErrorString(le));
    if (res->t_final) {
        if (rs == 8) memcpy(res->t_final, tf_host.data(), 8 * (size_t)N);
        else for (long long i = 0; i < N; ++i) res->t_final[i] = (double)((const float*)tf_host.data())[i];
    }
    if (res->ts && nslots > 0) {
        int k = 0;
        if (o->save_start != 0) res->ts[k++] = hp->t0;
        for (int i = 0; i < o->nsaveat && k < nslots; ++i) {
            const bool at_tf = (rs == 8) ? (o->saveat[i] == hp->tf) : ((float)o->saveat[i] == (float)hp->tf);
            if (at_tf && o->save_end == 0) continue;
            // the grid is stored in the program's real type
            res->ts[k++] = (rs == 8) ? o->saveat[i] : (double)(float)o->saveat[i];
        }
        if (k < nslots) res->ts[k++] = hp->tf;
    }
    float kms = 0, tms = 0;
    cudaEventElapsedTime(&kms, h->ev1, h->ev2);
    cudaEventElapsedTime(&tms, h->ev0, h->ev3);
    res->kernel_ms = kms; res->total_ms = tms;
    return B200ODE_OK;
}

int b200ode_reduce_sum_device(b200ode_handle h, int dtype, const void* x, int layout, int64_t count, int n,
                              double* out, void* stream) {
    if (!h || !x || !out) return fail(B200ODE_EINVAL, "NULL argument");
    if (n < 1 || count < 0) return fail(B200ODE_EINVAL, "bad n/count");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int nblocks = 2 * h->num_sms;
    CUDA_TRY(h->red_partial.ensure(sizeof(double) * (size_t)nblocks * n));
    long long ts = layout == B200ODE_LAYOUT_SOA ? 1 : n, cs = layout == B200ODE_LAYOUT_SOA ? count : 1;
    if (dtype == B200ODE_F32)
        k_reduce_partial<float><<<nblocks, 256, 0, s>>>((const float*)x, ts, cs, count, n, (double*)h->red_partial.ptr);
    else
        k_reduce_partial<double><<<nblocks, 256, 0, s>>>((const double*)x, ts, cs, count, n, (double*)h->red_partial.ptr);
    k_reduce_final<<<1, 256, 0, s>>>((const double*)h->red_partial.ptr, nblocks, n, out);
    CUDA_TRY(cudaGetLastError());
    return B200ODE_OK;
}

int b200ode_timeseries_meanvar_device(b200ode_handle h, int dtype, const void* us, int64_t count, int nslots, int n,
                                      double* mean, double* var, void* stream) {
    if (!h || !us || !mean) return fail(B200ODE_EINVAL, "NULL argument");
    if (count < 1 || nslots < 1 || n < 1) return fail(B200ODE_EINVAL, "count, nslots and n must be positive");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int cols = nslots * n;
    const int ychunks = (cols + 255) / 256;
    int nparts = (4 * h->num_sms + ychunks - 1) / ychunks;
    if ((long long)nparts > count) nparts = (int)count;
    CUDA_TRY(h->red_partial.ensure(sizeof(double) * (size_t)nparts * cols));
    double* part = (double*)h->red_partial.ptr;
    dim3 g(nparts, ychunks);
    if (dtype == B200ODE_F32) k_colsum_partial<float><<<g, 256, 0, s>>>((const float*)us, count, cols, nullptr, part);
    else k_colsum_partial<double><<<g, 256, 0, s>>>((const double*)us, count, cols, nullptr, part);
    k_colsum_final<<<ychunks, 256, 0, s>>>(part, nparts, cols, 1.0 / (double)count, mean);
    if (var) {
        if (dtype == B200ODE_F32) k_colsum_partial<float><<<g, 256, 0, s>>>((const float*)us, count, cols, mean, part);
        else k_colsum_partial<double><<<g, 256, 0, s>>>((const double*)us, count, cols, mean, part);
        k_colsum_final<<<ychunks, 256, 0, s>>>(part, nparts, cols, count > 1 ? 1.0 / (double)(count - 1) : 0.0, var);
    }
    CUDA_TRY(cudaGetLastError());
    return B200ODE_OK;
}

int b200ode_host_register(void* ptr, size_t bytes) {
    if (!ptr) return fail(B200ODE_EINVAL, "NULL pointer");
    CUDA_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return B200ODE_OK;
}
int b200ode_host_unregister(void* ptr) {
    if (!ptr) return fail(B200ODE_EINVAL, "NULL pointer");
    CUDA_TRY(cudaHostUnregister(ptr));
    return B200ODE_OK;
}

int b200ode_selftest_fastmath(b200ode_handle h, int64_t samples, uint64_t seed, int64_t* mismatches, int64_t* flagged) {
    if (!h || !mismatches) return fail(B200ODE_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    unsigned long long* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 8 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemsetAsync(d, 0, 8 * sizeof(unsigned long long), h->stream));
    k_selftest_fastmath<<<h->num_sms * 8, 256, 0, h->stream>>>((long long)samples, (unsigned long long)seed, d);
    unsigned long long out[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaError_t e = cudaMemcpyAsync(out, d, sizeof(out), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(B200ODE_ECUDA, std::string("selftest: ") + cudaGetErrorString(e));
    for (int k = 0; k < 6; ++k) mismatches[k] = (int64_t)out[k];
    if (flagged) *flagged = (int64_t)out[6];
    return B200ODE_OK;
}

int b200ode_measure_fma_peak(b200ode_handle h, int dtype, double* tflops, double* sm_clock_mhz) {
    if (!h || !tflops) return fail(B200ODE_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(h->scratch_t.ensure(64));
    const int iters = dtype == B200ODE_F32 ? 8192 : 4096;
    const int blocks = h->num_sms * 8, threads = 256;
    auto launch = [&]() {
        if (dtype == B200ODE_F32) k_fma_peak<float><<<blocks, threads, 0, h->stream>>>((float*)h->scratch_t.ptr, iters, 1.0000001f, 1e-7f);
        else k_fma_peak<double><<<blocks, threads, 0, h->stream>>>((double*)h->scratch_t.ptr, iters, 1.0000000001, 1e-10);
    };
    for (int w = 0; w < 3; ++w) launch();
    CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    const int reps = 5;
    for (int r = 0; r < reps; ++r) launch();
    CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    double flops = 2.0 * 64.0 * (double)iters * (double)blocks * threads * reps;
    *tflops = flops / (ms * 1e-3) / 1e12;
    if (sm_clock_mhz) {
        int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device);
        *sm_clock_mhz = khz / 1000.0;
    }
    return B200ODE_OK;
}

}  // extern "C"

// ===========================================================================
// Multi-GPU from ONE process (the reference analogue is EnsembleDistributed's pmap over workers,
// lib/DiffEqBase/test/downstream/distributed_ensemble.jl:43-51): trajectories are independent, so the ensemble is cut
// into contiguous chunks that are dealt round-robin to the devices (chunk c -> device c % ndev; several chunks per
// device so that a parameter sweep whose cost varies with the index is balanced), one host thread per device runs its
// chunks through the single-device entry points, and every chunk's outputs land directly in the caller's arrays at
// the chunk's offset — the "ordered gather" costs no extra copy.  No NCCL is needed: there is no exchange step in
// the integration; the ensemble mean is combined from per-chunk partial sums in chunk order (deterministic).
struct b200ode_multi_s {
    std::vector<b200ode_handle> dev;
};
struct b200ode_multi_program_s {
    b200ode_multi m = nullptr;
    std::vector<b200ode_program> prog;     // one per device
};

namespace {
const int kChunksPerDevice = 8;
struct ChunkPlan { long long c0, cn; int dev; };
std::vector<ChunkPlan> plan_chunks(long long N, int ndev) {
    std::vector<ChunkPlan> plan;
    if (N <= 0) return plan;
    long long nchunks = (long long)ndev * kChunksPerDevice;
    long long per = (N + nchunks - 1) / nchunks;
    per = ((per + 1023) / 1024) * 1024;                 // whole blocks of 1024 trajectories
    int c = 0;
    for (long long c0 = 0; c0 < N; c0 += per, ++c) plan.push_back({c0, std::min(per, N - c0), c % ndev});
    return plan;
}
size_t real_size(int dtype) { return dtype == B200ODE_F32 ? 4 : 8; }
}  // namespace

extern "C" {

int b200ode_multi_create(b200ode_multi* out, const int* device_ids, int ndev) {
    if (!out) return fail(B200ODE_EINVAL, "out is NULL");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return fail(B200ODE_ECUDA, "no CUDA device available; this library has no CPU fallback");
    if (ndev <= 0) ndev = count;                         // all visible devices
    if (ndev > count && !device_ids) return fail(B200ODE_EINVAL, "ndev exceeds the number of visible devices");
    b200ode_multi m = new b200ode_multi_s();
    for (int i = 0; i < ndev; ++i) {
        b200ode_handle h = nullptr;
        int rc = b200ode_create(&h, device_ids ? device_ids[i] : i);
        if (rc) { for (auto d : m->dev) b200ode_destroy(d); delete m; return rc; }
        m->dev.push_back(h);
    }
    *out = m;
    return B200ODE_OK;
}

int b200ode_multi_destroy(b200ode_multi m) {
    if (!m) return B200ODE_OK;
    for (auto d : m->dev) b200ode_destroy(d);
    delete m;
    return B200ODE_OK;
}

int b200ode_multi_device_count(b200ode_multi m) { return m ? (int)m->dev.size() : 0; }

int b200ode_multi_compile(b200ode_multi m, b200ode_multi_program* out, int alg, int dtype, int n, int np,
                          const char* rhs_src, const char* rhs_name, const char* jac_src, const char* jac_name,
                          const char* tgrad_src, const char* tgrad_name, const char* extra_options) {
    if (!m || !out) return fail(B200ODE_EINVAL, "multi handle/out is NULL");
    *out = nullptr;
    b200ode_multi_program mp = new b200ode_multi_program_s();
    mp->m = m;
    for (auto h : m->dev) {
        b200ode_program p = nullptr;
        int rc = b200ode_compile(h, &p, alg, dtype, n, np, rhs_src, rhs_name, jac_src, jac_name, tgrad_src, tgrad_name, extra_options);
        if (rc) { for (auto q : mp->prog) b200ode_program_destroy(q); delete mp; return rc; }
        mp->prog.push_back(p);
    }
    *out = mp;
    return B200ODE_OK;
}

int b200ode_multi_program_destroy(b200ode_multi_program mp) {
    if (!mp) return B200ODE_OK;
    for (auto q : mp->prog) b200ode_program_destroy(q);
    delete mp;
    return B200ODE_OK;
}

// shared driver of b200ode_multi_solve / b200ode_multi_reduce_mean
static int multi_run(b200ode_multi m, b200ode_multi_program mp, const B200Problem* hp, const B200Opts* o, B200Result* res,
                     double* mean) {
    if (!m || !mp || !hp || !o) return fail(B200ODE_EINVAL, "NULL argument");
    if (mp->m != m) return fail(B200ODE_EINVAL, "program was compiled for a different multi handle");
    const int ndev = (int)m->dev.size();
    const long long N = hp->trajectories;
    const b200ode_program p0 = mp->prog[0];
    const int n = p0->n, np = p0->np, nsave = p0->nsave;
    const size_t rs = real_size(p0->dtype);
    int rc = check_problem(N, hp->u0, hp->p, np, hp->t0, hp->tf, o, p0->reverse);
    if (rc) return rc;
    if (!mean && (!res || !res->u_final)) return fail(B200ODE_EINVAL, "result.u_final is required");
    const int nslots = nslots_typed(hp, o, p0->dtype);
    const std::vector<ChunkPlan> plan = plan_chunks(N, ndev);
    std::vector<std::vector<double>> partial(plan.size());          // per-chunk sums for the ensemble mean
    std::vector<int> status(ndev, 0);
    std::vector<std::string> message(ndev);
    std::vector<double> kms(ndev, 0.0), tms(ndev, 0.0);
    auto worker = [&](int d) {
        b200ode_handle h = m->dev[d];
        b200ode_program prog = mp->prog[d];
        std::vector<double> ts_scratch((size_t)std::max(nslots, 1));
        std::vector<char> uf_scratch;
        for (size_t ci = 0; ci < plan.size(); ++ci) {
            if (plan[ci].dev != d) continue;
            const long long c0 = plan[ci].c0, cn = plan[ci].cn;
            B200Problem sub = *hp;
            sub.trajectories = cn;
            sub.u0 = hp->u0_shared ? hp->u0 : (const char*)hp->u0 + rs * n * (size_t)c0;
            sub.p = (np > 0 && !hp->p_shared) ? (const void*)((const char*)hp->p + rs * np * (size_t)c0) : hp->p;
            sub.tspans = hp->tspans ? hp->tspans + 2 * (size_t)c0 : nullptr;
            B200Result r{};
            if (res) {
                r = *res;
                auto off = [&](void* base, size_t stride) -> void* { return base ? (char*)base + stride * (size_t)c0 : nullptr; };
                r.u_final = off(res->u_final, rs * n);
                r.t_final = (double*)off(res->t_final, sizeof(double));
                r.us = off(res->us, rs * (size_t)nsave * (size_t)nslots);
                r.ts = (ci == 0) ? res->ts : (res->ts ? ts_scratch.data() : nullptr);   // one writer for the shared grid
                r.nsaved = (int32_t*)off(res->nsaved, 4); r.naccept = (int32_t*)off(res->naccept, 4);
                r.nreject = (int32_t*)off(res->nreject, 4); r.nf = (int32_t*)off(res->nf, 4);
                r.njacs = (int32_t*)off(res->njacs, 4); r.nw = (int32_t*)off(res->nw, 4);
                r.nsolve = (int32_t*)off(res->nsolve, 4); r.retcode = (int32_t*)off(res->retcode, 4);
            }
            if (!r.u_final) { uf_scratch.resize(rs * n * (size_t)cn); r.u_final = uf_scratch.data(); }
            int st = b200ode_solve(h, prog, &sub, o, &r);
            if (st) { status[d] = st; message[d] = g_last_error; return; }
            kms[d] += r.kernel_ms; tms[d] += r.total_ms;
            if (mean) {
                // the chunk's final states are still resident in the handle's device buffer: reduce them there
                partial[ci].assign(n, 0.0);
                cudaSetDevice(h->device);
                if (h->stat_out.ensure(sizeof(double) * n) != cudaSuccess) { status[d] = B200ODE_ECUDA; message[d] = "out of memory"; return; }
                st = b200ode_reduce_sum_device(h, prog->dtype, h->out_uf.ptr, B200ODE_LAYOUT_AOS, cn, n, (double*)h->stat_out.ptr, h->stream);
                if (st == 0 && cudaMemcpyAsync(partial[ci].data(), h->stat_out.ptr, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) st = B200ODE_ECUDA;
                if (st == 0 && cudaStreamSynchronize(h->stream) != cudaSuccess) st = B200ODE_ECUDA;
                if (st) { status[d] = st; message[d] = g_last_error; return; }
            }
        }
    };
    std::vector<std::thread> threads;
    for (int d = 1; d < ndev; ++d) threads.emplace_back(worker, d);
    worker(0);
    for (auto& t : threads) t.join();
    for (int d = 0; d < ndev; ++d)
        if (status[d]) return fail(status[d], "device " + std::to_string(m->dev[d]->device) + ": " + message[d]);
    if (res) {
        res->kernel_ms = *std::max_element(kms.begin(), kms.end());
        res->total_ms = *std::max_element(tms.begin(), tms.end());
    }
    if (mean) {
        for (int c = 0; c < n; ++c) {
            double acc = 0.0;
            for (size_t ci = 0; ci < plan.size(); ++ci) acc += partial[ci][c];      // chunk order: deterministic
            mean[c] = N > 0 ? acc / (double)N : 0.0;
        }
    }
    return B200ODE_OK;
}

int b200ode_multi_solve(b200ode_multi m, b200ode_multi_program mp, const B200Problem* hp, const B200Opts* o, B200Result* res) {
    if (mp && !mp->prog.empty() && mp->prog[0]->everystep)
        return fail(B200ODE_EUNSUPPORTED, "save_everystep programs are single-device (ragged output)");
    return multi_run(m, mp, hp, o, res, nullptr);
}

int b200ode_multi_reduce_mean(b200ode_multi m, b200ode_multi_program mp, const B200Problem* hp, const B200Opts* o,
                              double* mean, B200Result* res) {
    if (!mean) return fail(B200ODE_EINVAL, "mean is NULL");
    if (mp && !mp->prog.empty() && mp->prog[0]->everystep)
        return fail(B200ODE_EUNSUPPORTED, "save_everystep programs are single-device (ragged output)");
    if (o && o->saveat && o->nsaveat > 0) return fail(B200ODE_EINVAL, "b200ode_multi_reduce_mean reduces final states: no saveat grid");
    return multi_run(m, mp, hp, o, res, mean);
}

}  // extern "C"
