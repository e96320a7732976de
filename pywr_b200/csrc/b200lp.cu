// b200lp.cu — C-ABI (include/b200lp.h) of the B200 scenario-batched LP engine: handle management,
// device memory, host<->device copies and kernel dispatch.  No torch, no Python, no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200lp.h"
#include "lp_kernel.cuh"
#include "lp_kernel_cta.cuh"
#include "pdhg_engine.cuh"
#include "rsx_kernels.cuh"

using namespace b200lp;

static thread_local std::string g_last_error;

static int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(B200LP_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));              \
    } while (0)


// Opt-in to more than 48 KB of dynamic shared memory, per device and per kernel, raise-only: the attribute is a property of
// (function, device), engines of different sizes share kernels, and a lower value set for a later engine would break the
// launches of a larger one that is still alive.
static cudaError_t ensure_smem(const void *func, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> high;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    size_t &hw = high[std::make_pair(dev, func)];
    if (bytes <= hw) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) hw = bytes;
    return e;
}

typedef void (*launch_fn)(const DevStructure &, const DevState &, int S, double days, const double *slots,
                          double *node_flow, double *col_flow, double *objective, cudaStream_t stream, int block);

template <int G, int RPL, int NC>
static void launch_variant(const DevStructure &st, const DevState &state, int S, double days, const double *slots,
                           double *node_flow, double *col_flow, double *objective, cudaStream_t stream, int block) {
    const int groups_per_block = block / G;
    const int grid = (S + groups_per_block - 1) / groups_per_block;
    const size_t gbytes = (GroupSmem<G, RPL, NC>::bytes(st.n_val) + 15) & ~(size_t)15;
    const size_t smem = gbytes * groups_per_block;
    ensure_smem((const void *)lp_step_kernel<G, RPL, NC>, smem);  // a failure surfaces as the launch error
    lp_step_kernel<G, RPL, NC><<<grid, block, smem, stream>>>(st, state, S, days, slots, node_flow, col_flow, objective);
}

typedef void (*run_fn)(const DevStructure &, const DevState &, const DevProgram &, int S, int t0, int nsteps,
                       cudaStream_t stream, int block);

// Blocks per SM of the run kernel (lp_kernel.cuh::lp_run_kernel MINB) for this batch.  The kernel is a latency-bound chain
// per scenario, so a launch takes (number of waves) x (chain time): 4,096 thames scenarios are 512 blocks, two waves of the
// 444 resident blocks of the 168-register build but ONE wave of the 592 of the 128-register build (measured on B200: 66.9 ->
// 38.2 ms per 30-year run), while 1,024 two_reservoir scenarios fit one wave either way and the 168-register chain is 25 %
// shorter (34.2 vs 42.7 ms).  B200LP_RUN_BLOCKS=3|4 overrides.
template <int RPL, int NC>
static int pick_run_blocks(int grid) {
    if (RPL != 1) return 3;
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    const char *v = getenv("B200LP_RUN_BLOCKS");
    const int forced = v ? atoi(v) : 0;
    if (forced == 3 || forced == 4) return forced;
    // up to two waves the blocks run in lockstep (whole waves count); beyond that the scheduler refills SMs as blocks finish
    // (fractional waves) and the spilled values also cost issue slots.  Measured on B200 (ms per run, 3 / 4 blocks per SM):
    // thames 4,096: 66.9 / 38.2; thames 65,536 x 3 y: 64.7 / 53.1; two_reservoir 2,048: 70.0 / 52.4; 32,768 x 5 y: 70.6 / 76.8
    const double w3 = (double)grid / (3.0 * sms), w4 = (double)grid / (4.0 * sms);
    if (w3 <= 1.0) return 3;
    const bool lockstep = w3 <= 2.0;
    const double waves3 = lockstep ? ceil(w3) : w3, waves4 = lockstep ? ceil(w4) : w4;
    const double slowdown = NC <= 8 ? 1.05 : (NC <= 13 ? (lockstep ? 1.25 : 1.45) : 1.6);  // chain time, 128- / 168-register build
    return waves4 * slowdown < waves3 ? 4 : 3;
}

template <int G, int RPL, int NC, int MINB>
static void run_variant_mb(const DevStructure &st, const DevState &state, const DevProgram &pg, int S, int t0, int nsteps,
                           cudaStream_t stream, int block, int grid, size_t smem) {
    ensure_smem((const void *)lp_run_kernel<G, RPL, NC, MINB>, smem);
    lp_run_kernel<G, RPL, NC, MINB><<<grid, block, smem, stream>>>(st, state, pg, S, t0, nsteps);
}

template <int G, int RPL, int NC>
static void run_variant(const DevStructure &st, const DevState &state, const DevProgram &pg, int S, int t0, int nsteps,
                        cudaStream_t stream, int block) {
    const int groups_per_block = block / G;
    const int grid = (S + groups_per_block - 1) / groups_per_block;
    const size_t gbytes = (GroupSmem<G, RPL, NC>::bytes(st.n_val) + 15) & ~(size_t)15;
    const size_t smem = gbytes * groups_per_block + sizeof(int) * (size_t)(((pg.n_code + 3) & ~3) + pg.n_fop_levels * 32 * FOP_WORDS + 4);
    if (RPL == 1 && block == 128 && pick_run_blocks<RPL, NC>(grid) == 4)
        run_variant_mb<G, RPL, NC, (RPL == 1 ? 4 : 1)>(st, state, pg, S, t0, nsteps, stream, block, grid, smem);
    else
        run_variant_mb<G, RPL, NC, (RPL == 1 ? B200LP_RUN_MIN_BLOCKS : 1)>(st, state, pg, S, t0, nsteps, stream, block, grid, smem);
}

struct Variant {
    int G, RPL, NC;
    launch_fn launch;
    run_fn run;
    size_t (*smem_per_group)(int n_slots);
};

#define VARIANT(G, RPL, NC) \
    { G, RPL, NC, &launch_variant<G, RPL, NC>, &run_variant<G, RPL, NC>, &GroupSmem<G, RPL, NC>::bytes }

// ordered from the cheapest; the first one that fits (rows <= G*RPL, cols <= NC) is used unless
// B200LP_VARIANT forces another one that fits.  Variants 8.. pack several scenarios into one warp
// (G < 32 lanes per scenario, more rows per lane): fewer, fatter warps for latency-bound batches.
static const Variant kVariants[] = {
    VARIANT(8, 1, 4),   VARIANT(16, 1, 8),  VARIANT(32, 1, 8),  VARIANT(32, 1, 13),
    VARIANT(32, 1, 16), VARIANT(32, 1, 24), VARIANT(32, 2, 16), VARIANT(32, 2, 32),
    VARIANT(16, 2, 13), VARIANT(8, 4, 13),  VARIANT(8, 2, 8),   VARIANT(4, 4, 8),
};
static const int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

// Host copy of the LP structure with every ref resolved to an index into the per-scenario value vector
//     V = [ slots | constants | storage volumes | storage pc | K_* built-ins ]
// (what DevStructure holds as device arrays).  Pure host code: b200lp_create uploads it, b200lp_load_program pre-decodes
// the ops against it and jit_run.inc generates the structure-specialised run kernel from it.
struct HostStructure {
    int m = 0, n = 0, n_slots = 0, n_consts = 0, n_storage = 0, n_dyn_a = 0, cost_dynamic = 0, rows_cap = 0;
    int vol_base = 0, pc_base = 0, n_val = 0, k_base = 0;
    std::vector<int> row_class, row_a, row_b, st_kind, st_row, st_vol, st_minv, st_maxv, st_act;
    int cost_clip = 0;
    std::vector<int> cost_indptr, cost_idx;   // objective: c_j = sum cost_w[k] * V[cost_idx[k]] (a3)
    std::vector<double> cost_w;
};

static bool resolve_structure(const b200lp_structure *s, int rows_cap, HostStructure &hs, std::string &err) {
    const int m = s->n_rows, n = s->n_cols;
    hs.m = m; hs.n = n; hs.n_slots = s->n_slots; hs.n_consts = s->n_consts; hs.n_storage = s->n_storage;
    hs.n_dyn_a = s->n_dyn_a; hs.rows_cap = rows_cap;
    hs.cost_dynamic = 0;
    for (int k = 0; k < s->cost_indptr[n]; k++)
        if (s->cost_ref[k] >= 0) hs.cost_dynamic = 1;
    hs.cost_clip = s->cost_clip;
    hs.cost_indptr.assign(s->cost_indptr, s->cost_indptr + n + 1);
    hs.cost_w.assign(s->cost_w, s->cost_w + s->cost_indptr[n]);
    hs.cost_idx.resize(s->cost_indptr[n]);
    for (int k = 0; k < s->cost_indptr[n]; k++) {
        const int ref = s->cost_ref[k];
        if (ref >= 0 ? ref >= s->n_slots : (ref == B200LP_REF_STATE || (-ref - 1) >= s->n_consts)) { err = "cost ref out of range"; return false; }
        hs.cost_idx[k] = ref >= 0 ? ref : s->n_slots + (-ref - 1);
    }
    hs.vol_base = s->n_slots + s->n_consts;
    hs.pc_base = hs.vol_base + s->n_storage;
    hs.n_val = hs.pc_base + s->n_storage;
    hs.k_base = hs.n_val;
    const int n_slots = s->n_slots, n_consts = s->n_consts;
    auto ref_ok = [&](int ref) { return ref >= 0 ? ref < n_slots : (ref != B200LP_REF_STATE && (-ref - 1) < n_consts); };
    auto vidx = [&](int ref) { return ref >= 0 ? ref : n_slots + (-ref - 1); };
    hs.row_class.assign(rows_cap, ROWC_PAD); hs.row_a.assign(rows_cap, 0); hs.row_b.assign(rows_cap, 0);
    hs.st_row.assign(s->n_storage ? s->n_storage : 1, 0);
    for (int i = 0; i < m; i++) {
        if (s->row_kind[i] == B200LP_ROW_FLOW) {
            if (!ref_ok(s->row_lb_ref[i]) || !ref_ok(s->row_ub_ref[i])) { err = "row ref out of range"; return false; }
            hs.row_class[i] = (s->row_lb_ref[i] < 0 && s->row_ub_ref[i] < 0) ? ROWC_STATIC : ROWC_FLOW;
            hs.row_a[i] = vidx(s->row_lb_ref[i]); hs.row_b[i] = vidx(s->row_ub_ref[i]);
        } else {
            if (s->row_lb_ref[i] < 0 || s->row_lb_ref[i] >= s->n_storage) { err = "storage index out of range"; return false; }
            hs.row_class[i] = ROWC_STORAGE; hs.row_a[i] = s->row_lb_ref[i];
            hs.st_row[s->row_lb_ref[i]] = i;
        }
        for (int k = s->a_indptr[i]; k < s->a_indptr[i + 1]; k++)
            if (s->a_col[k] < 0 || s->a_col[k] >= n) { err = "column index out of range"; return false; }
    }
    hs.st_kind.assign(s->st_kind, s->st_kind + s->n_storage);
    hs.st_vol.resize(s->n_storage); hs.st_minv.resize(s->n_storage); hs.st_maxv.resize(s->n_storage); hs.st_act.resize(s->n_storage);
    for (int k = 0; k < s->n_storage; k++) {
        if (s->st_vol_ref[k] != B200LP_REF_STATE && !ref_ok(s->st_vol_ref[k])) { err = "storage volume ref out of range"; return false; }
        if (!ref_ok(s->st_minv_ref[k]) || !ref_ok(s->st_maxv_ref[k]) || !ref_ok(s->st_active_ref[k])) { err = "storage ref out of range"; return false; }
        hs.st_vol[k] = s->st_vol_ref[k] == B200LP_REF_STATE ? hs.vol_base + k : vidx(s->st_vol_ref[k]);
        hs.st_minv[k] = vidx(s->st_minv_ref[k]); hs.st_maxv[k] = vidx(s->st_maxv_ref[k]); hs.st_act[k] = vidx(s->st_active_ref[k]);
    }
    return true;
}

// Ops of a program with every operand resolved to an index into V (decode_program below).
struct DecodedProgram {
    std::vector<int> tab_dst, tab_table, tab_offset, code, rec_idx, rec_src;
    bool has_lag = false;   // some op reads a recorder's sample of the previous timestep (B200LP_OP_LAG)
};

namespace jit {
// what the generator of the structure-specialised run kernel (jit_run.inc) needs, all host copies
struct Spec {
    int G = 0, NC = 0, minb = 3;
    HostStructure hs;
    std::vector<double> consts;
    std::vector<int> st_kind;
    DecodedProgram dec;
    int n_recorders = 0;
    std::vector<int> rec_kind, rec_row, rec_indptr, rec_col;
    std::vector<double> rec_factor, rec_w;
    std::vector<int> stflow_indptr, stflow_col;
    std::vector<double> stflow_w;
    std::vector<int> roll_len, roll_off;
    double uniform_days = 0.0;  // > 0: every timestep of the program has this length
    bool defer_recorders = false;  // recorder samples of timestep t taken at the top of t + 1 (latency-bound batches)
    bool valid = false;
};
}  // namespace jit

struct b200lp_handle {
    int device = 0;
    int S = 0;
    int variant = -1;
    bool use_cta = false;  // one CTA per scenario, dictionary in shared memory (lp_kernel_cta.cuh)
    size_t cta_smem = 0;
    int block = 128;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    DevStructure ds{};
    DevState state{};
    std::vector<void *> owned;  // device allocations
    double *d_consts = nullptr;
    double *d_slots = nullptr, *d_node_flow = nullptr, *d_col_flow = nullptr, *d_objective = nullptr;
    int *h_fail = nullptr;  // pinned [2]
    unsigned long long *h_counters = nullptr;  // pinned [2]
    int nnz = 0;
    int last_first_bad = -1, last_code = 0;
    b200lp_stats stats{};
    // device-resident run
    bool has_program = false;
    DevProgram dp{};
    std::vector<void *> prog_owned;
    std::vector<double *> rec_ptr;   // device outputs
    std::vector<int> rec_kind;
    bool agg_only = false;           // B200LP_PROG_AGG_ONLY: array recorders keep aggregates only
    std::vector<double> init_vol, init_pc;
    std::vector<int> roll_len, roll_off;   // rolling virtual storages: ring-buffer rows per storage
    std::vector<double> roll_init;
    std::vector<int> table_rows, table_cols;
    std::vector<long long> table_offset;
    int prog_steps = 0;
    cudaStream_t s_in = nullptr, s_out = nullptr;  // copy streams of b200lp_run_pipelined
    std::vector<cudaEvent_t> pipe_events;
    std::vector<int> h_st_maxv_idx;  // V index of every storage's max_volume (host copy, for op pre-decoding)
    HostStructure hs;                // resolved LP structure (register / CTA engines)
    jit::Spec jspec;                 // what the loaded program's specialised kernel is generated from (jit_run.inc)
    cudaKernel_t jit_kernel = nullptr;
    bool jit_stale = false;          // constants changed since the kernel was generated
    std::string jit_note;            // why the generic run kernel is used instead
    std::vector<double> consts_host;
    std::vector<int> st_kind_host;
    struct PdhgEngine *pdhg = nullptr;  // large-LP path (pdhg_engine.cuh, rsx_kernels.cuh): LPs beyond the dictionary kernels
    struct PlaneHost *plane = nullptr;  // device-resident program on the large-LP path (plane_program.inc)
    std::vector<double> ts_days_host;
};

template <typename T>
static cudaError_t upload(b200lp_handle *h, const T *src, size_t count, const T **out) {
    T *p = nullptr;
    const size_t bytes = sizeof(T) * (count ? count : 1);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return e;
    h->owned.push_back(p);
    e = cudaMemset(p, 0, bytes);
    if (e != cudaSuccess) return e;
    if (count && src) e = cudaMemcpy(p, src, sizeof(T) * count, cudaMemcpyHostToDevice);
    *out = p;
    return e;
}

template <typename T>
static cudaError_t dalloc(b200lp_handle *h, size_t count, T **out) {
    T *p = nullptr;
    const size_t bytes = sizeof(T) * (count ? count : 1);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return e;
    h->owned.push_back(p);
    *out = p;
    return cudaMemset(p, 0, bytes);
}

// running aggregates [recorder][3][S]: sum = 0, min = +inf, max = -inf
__global__ void agg_reset_kernel(double *p, size_t n, size_t S) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t which = (i / S) % 3;
    p[i] = which == 0 ? 0.0 : (which == 1 ? INFINITY : -INFINITY);
}

__global__ void fill_kernel(double *p, size_t n, double v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

static void free_program(b200lp_handle *h);

#include "plane_kernels.cuh"
#include "pdhg_host.inc"
#include "plane_program.inc"

extern "C" {

const char *b200lp_last_error(void) { return g_last_error.c_str(); }
int b200lp_abi_version(void) { return B200LP_ABI_VERSION; }

int b200lp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int b200lp_create(const b200lp_structure *s, int32_t n_scenarios, int32_t device, b200lp_handle **out) {
    return b200lp_create_ex(s, n_scenarios, device, 0u, out);
}

int b200lp_create_ex(const b200lp_structure *s, int32_t n_scenarios, int32_t device, uint32_t flags, b200lp_handle **out) {
    if (!s || !out) return fail(B200LP_E_INVALID, "null argument");
    if (s->abi_version != B200LP_ABI_VERSION) return fail(B200LP_E_INVALID, "ABI version mismatch");
    if (n_scenarios <= 0 || s->n_rows <= 0 || s->n_cols <= 0) return fail(B200LP_E_INVALID, "empty problem");
    int ndev = b200lp_device_count();
    if (ndev <= 0) return fail(B200LP_E_NO_DEVICE, "no CUDA device visible: this engine has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(B200LP_E_INVALID, "device index out of range");
    const int m = s->n_rows, n = s->n_cols, N = s->n_nodes;
    int variant = -1;
    for (int v = 0; v < kNumVariants; v++)
        if (m <= kVariants[v].G * kVariants[v].RPL && n <= kVariants[v].NC) { variant = v; break; }
    // Small batches are latency bound (one dependent chain per scenario and timestep): a whole warp per scenario keeps a
    // scenario that pivots from stalling the one it would share a warp with.  Measured on B200 (thames, specialised run
    // kernel, M scenario-timesteps/s, 16 / 32 lanes per scenario): 512 scenarios 403 / 629, 1,024: 807 / 1,108, 2,048: 1,475 /
    // 1,584, 4,096: 2,317 / 1,800 (profiles/r2_sweep_variants.log).
    if (variant >= 0 && kVariants[variant].RPL == 1 && kVariants[variant].G < 32 && n_scenarios <= 2048)
        for (int v = 0; v < kNumVariants; v++)
            if (kVariants[v].G == 32 && kVariants[v].RPL == 1 && m <= 32 && n <= kVariants[v].NC) { variant = v; break; }
    if (const char *force = getenv("B200LP_VARIANT")) {
        int v = atoi(force);
        if (v >= 0 && v < kNumVariants && m <= kVariants[v].G * kVariants[v].RPL && n <= kVariants[v].NC) variant = v;
    }
    if (getenv("B200LP_FORCE_PDHG")) return create_pdhg(s, n_scenarios, device, out);
    constexpr int CTA_THREADS = 128;
    bool use_cta = false;
    size_t cta_smem = 0;
    if (variant < 0 || getenv("B200LP_FORCE_CTA")) {
        const int n_val = s->n_slots + s->n_consts + 2 * s->n_storage;
        cta_smem = CtaDict<CTA_THREADS>::bytes(m, n, n_val);
        // beyond shared memory: the large-LP path (pdhg_engine.cuh front end; revised simplex or PDHG iterations); also for
        // mid-size LPs when the handle is meant for device-resident programs, which the CTA-per-scenario engine does not run
        if (cta_smem > 227 * 1024) return create_pdhg(s, n_scenarios, device, out);
        if ((flags & B200LP_CREATE_FOR_PROGRAM) && s->n_dyn_a == 0 && !getenv("B200LP_FORCE_CTA")) return create_pdhg(s, n_scenarios, device, out);
        use_cta = true;
        variant = 0;
    }
    CUDA_TRY(cudaSetDevice(device));
    b200lp_handle *h = new b200lp_handle();
    h->device = device; h->S = n_scenarios; h->variant = variant;
    h->use_cta = use_cta; h->cta_smem = cta_smem;
    if (const char *b = getenv("B200LP_BLOCK")) {
        int v = atoi(b);
        if (v >= 32 && v <= 128 && v % 32 == 0) h->block = v;
    }
    const Variant &V = kVariants[variant];
    const int rows_cap = use_cta ? m : V.G * V.RPL, nc = use_cta ? n : V.NC;
    const size_t S = (size_t)n_scenarios;
    h->nnz = s->a_indptr[m];

    DevStructure &ds = h->ds;
    ds.m = m; ds.n = n; ds.N = N; ds.n_slots = s->n_slots; ds.n_consts = s->n_consts;
    ds.n_storage = s->n_storage; ds.n_dyn_a = s->n_dyn_a; ds.cost_clip = s->cost_clip;
    ds.rows_cap = rows_cap; ds.nc = nc;
    ds.cost_dynamic = 0;
    for (int k = 0; k < s->cost_indptr[n]; k++)
        if (s->cost_ref[k] >= 0) ds.cost_dynamic = 1;

    HostStructure &hs = h->hs;
    {
        std::string err;
        if (!resolve_structure(s, rows_cap, hs, err)) { delete h; return fail(B200LP_E_INVALID, err); }
    }
    ds.vol_base = hs.vol_base; ds.pc_base = hs.pc_base; ds.n_val = hs.n_val; ds.k_base = hs.k_base;
    const int n_slots = s->n_slots, n_consts = s->n_consts;
    auto ref_ok = [&](int ref) { return ref >= 0 ? ref < n_slots : (ref != B200LP_REF_STATE && (-ref - 1) < n_consts); };
    auto vidx = [&](int ref) { return ref >= 0 ? ref : n_slots + (-ref - 1); };
    const std::vector<int> &row_class = hs.row_class, &row_a = hs.row_a, &row_b = hs.row_b, &st_row = hs.st_row;
    const std::vector<int> &st_vol = hs.st_vol, &st_minv = hs.st_minv, &st_maxv = hs.st_maxv, &st_act = hs.st_act;
    h->h_st_maxv_idx = st_maxv;
    h->consts_host.assign(s->consts, s->consts + s->n_consts);
    h->st_kind_host.assign(s->st_kind, s->st_kind + s->n_storage);
    const int ncost = s->cost_indptr[n];
    std::vector<int> cost_idx(ncost), dyn_idx(s->n_dyn_a);
    for (int k = 0; k < ncost; k++) {
        if (!ref_ok(s->cost_ref[k])) { delete h; return fail(B200LP_E_INVALID, "cost ref out of range"); }
        cost_idx[k] = vidx(s->cost_ref[k]);
    }
    for (int k = 0; k < s->n_dyn_a; k++) {
        if (!ref_ok(s->dyn_a_ref[k])) { delete h; return fail(B200LP_E_INVALID, "coefficient ref out of range"); }
        dyn_idx[k] = vidx(s->dyn_a_ref[k]);
    }

    std::vector<double> A((size_t)nc * rows_cap, 0.0);
    for (int i = 0; i < m; i++)
        for (int k = s->a_indptr[i]; k < s->a_indptr[i + 1]; k++) A[(size_t)s->a_col[k] * rows_cap + i] += s->a_val[k];
    std::vector<int> dyn_row(s->n_dyn_a), dyn_col(s->n_dyn_a);
    std::vector<double> dyn_static(s->n_dyn_a);
    for (int k = 0; k < s->n_dyn_a; k++) {
        const int nz = s->dyn_a_nz[k];
        int row = 0;
        while (s->a_indptr[row + 1] <= nz) row++;
        dyn_row[k] = row; dyn_col[k] = s->a_col[nz]; dyn_static[k] = s->a_val[nz];
    }

#define UP(field, src, count)                                                     \
    do {                                                                          \
        cudaError_t _e = upload(h, src, (size_t)(count), &ds.field);              \
        if (_e != cudaSuccess) {                                                  \
            b200lp_destroy(h);                                                    \
            return fail(B200LP_E_CUDA, std::string("upload " #field ": ") + cudaGetErrorString(_e)); \
        }                                                                         \
    } while (0)
    UP(A_dense, A.data(), A.size());
    UP(row_class, row_class.data(), rows_cap);
    UP(row_a, row_a.data(), rows_cap);
    UP(row_b, row_b.data(), rows_cap);
    UP(st_kind, s->st_kind, s->n_storage);
    UP(st_vol_idx, st_vol.data(), s->n_storage);
    UP(st_minv_idx, st_minv.data(), s->n_storage);
    UP(st_maxv_idx, st_maxv.data(), s->n_storage);
    UP(st_active_idx, st_act.data(), s->n_storage);
    UP(st_row, st_row.data(), s->n_storage);
    UP(cost_indptr, s->cost_indptr, n + 1);
    UP(cost_idx, cost_idx.data(), ncost);
    UP(cost_w, s->cost_w, ncost);
    UP(dyn_a_row, dyn_row.data(), s->n_dyn_a);
    UP(dyn_a_col, dyn_col.data(), s->n_dyn_a);
    UP(dyn_a_idx, dyn_idx.data(), s->n_dyn_a);
    UP(dyn_a_scale, s->dyn_a_scale, s->n_dyn_a);
    UP(dyn_a_static, dyn_static.data(), s->n_dyn_a);
    UP(flow_indptr, s->flow_indptr, N + 1);
    UP(flow_col, s->flow_col, s->flow_indptr[N]);
    UP(flow_w, s->flow_w, s->flow_indptr[N]);
    UP(consts, s->consts, s->n_consts);
#undef UP
    h->d_consts = const_cast<double *>(ds.consts);

#define DA(ptr, count)                                                           \
    do {                                                                         \
        cudaError_t _e = dalloc(h, (size_t)(count), &ptr);                       \
        if (_e != cudaSuccess) {                                                 \
            b200lp_destroy(h);                                                   \
            return fail(B200LP_E_CUDA, std::string("alloc " #ptr ": ") + cudaGetErrorString(_e)); \
        }                                                                        \
    } while (0)
    DA(h->state.T, S * nc * rows_cap);
    DA(h->state.d, S * nc);
    DA(h->state.head, S * rows_cap);
    DA(h->state.nb, S * nc);
    DA(h->state.nbstat, S * nc);
    DA(h->state.meta, S * 4);
    DA(h->state.vol, S * (s->n_storage ? s->n_storage : 1));
    DA(h->state.counters, 2);
    DA(h->state.fail, 2);
    DA(h->d_slots, S * (s->n_slots ? s->n_slots : 1));
    DA(h->d_node_flow, S * N);
    DA(h->d_col_flow, S * n);
    DA(h->d_objective, S);
#undef DA
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess ||
        cudaHostAlloc((void **)&h->h_fail, 2 * sizeof(int), cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc((void **)&h->h_counters, 2 * sizeof(unsigned long long), cudaHostAllocDefault) != cudaSuccess) {
        b200lp_destroy(h);
        return fail(B200LP_E_CUDA, "stream / event / pinned allocation failed");
    }
    h->stats.n_rows = m; h->stats.n_cols = n; h->stats.n_nonzero = h->nnz; h->stats.kernel_variant = use_cta ? 100 : variant;
    *out = h;
    int rc = b200lp_reset(h, nullptr);
    if (rc != 0) { b200lp_destroy(h); *out = nullptr; return rc; }
    return B200LP_OK;
}

void b200lp_destroy(b200lp_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    free_program(h);
    pdhg_destroy(h);
    for (void *p : h->owned) cudaFree(p);
    if (h->h_fail) cudaFreeHost(h->h_fail);
    if (h->h_counters) cudaFreeHost(h->h_counters);
    for (cudaEvent_t e : h->pipe_events) cudaEventDestroy(e);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int b200lp_reset(b200lp_handle *h, const double *initial_volume) {
    if (!h) return fail(B200LP_E_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->pdhg) return pdhg_reset(h, initial_volume);
    const int threads = 128;
    lp_reset_kernel<<<(h->S + threads - 1) / threads, threads, 0, h->stream>>>(h->state, h->S, h->ds.m, h->ds.n,
                                                                             h->ds.rows_cap, h->ds.nc);
    h->stats.kernel_launches++;
    CUDA_TRY(cudaGetLastError());
    if (initial_volume && h->ds.n_storage > 0) {
        const size_t bytes = sizeof(double) * (size_t)h->ds.n_storage * h->S;
        CUDA_TRY(cudaMemcpyAsync(h->state.vol, initial_volume, bytes, cudaMemcpyHostToDevice, h->stream));
        h->stats.h2d_bytes += (double)bytes;
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->last_first_bad = -1; h->last_code = 0;
    return B200LP_OK;
}

int b200lp_set_consts(b200lp_handle *h, const double *consts) {
    if (!h || !consts) return fail(B200LP_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->pdhg) {
        const int rc = pdhg_set_consts(h, consts);
        if (rc == 0 && h->plane) {
            plane_fill_consts<<<(h->S + 127) / 128, 128, 0, h->stream>>>(h->plane->Q, h->d_consts);
            CUDA_TRY(cudaStreamSynchronize(h->stream));
        }
        return rc;
    }
    if (h->ds.n_consts > 0) {
        CUDA_TRY(cudaMemcpyAsync(h->d_consts, consts, sizeof(double) * h->ds.n_consts, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        h->consts_host.assign(consts, consts + h->ds.n_consts);
        if (h->has_program && h->jspec.valid) {  // the specialised kernel holds the constants as immediates
            h->jspec.consts = h->consts_host;
            h->jit_stale = true;
        }
    }
    return B200LP_OK;
}

static int launch_step(b200lp_handle *h, double days, const double *d_slots, double *d_node_flow, double *d_col_flow,
                       double *d_objective) {
    if (h->use_cta) {
        ensure_smem((const void *)lp_step_kernel_cta<128>, h->cta_smem);
        lp_step_kernel_cta<128><<<h->S, 128, h->cta_smem, h->stream>>>(h->ds, h->state, h->S, days, d_slots, d_node_flow,
                                                                      d_col_flow, d_objective);
    } else {
        const Variant &V = kVariants[h->variant];
        V.launch(h->ds, h->state, h->S, days, d_slots, d_node_flow, d_col_flow, d_objective, h->stream, h->block);
    }
    h->stats.kernel_launches++;
    h->stats.steps++;
    h->stats.scenario_steps += h->S;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(B200LP_E_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
    return B200LP_OK;
}

int b200lp_step_device(b200lp_handle *h, double days, const double *d_slots, double *d_node_flow, double *d_col_flow,
                       double *d_objective) {
    if (!h) return fail(B200LP_E_INVALID, "null handle");
    if (!(days > 0.0)) return fail(B200LP_E_INVALID, "days must be positive");
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->pdhg)  // iterates to convergence: synchronises with the host between batches of iterations
        return pdhg_step_device(h, days, d_slots ? d_slots : h->d_slots, d_node_flow, d_col_flow, d_objective, nullptr);
    return launch_step(h, days, d_slots ? d_slots : h->d_slots, d_node_flow, d_col_flow, d_objective);
}

int b200lp_step(b200lp_handle *h, double days, const double *slots, double *node_flow, double *col_flow,
                double *objective) {
    if (!h) return fail(B200LP_E_INVALID, "null handle");
    if (!(days > 0.0)) return fail(B200LP_E_INVALID, "days must be positive");
    if (h->ds.n_slots > 0 && !slots) return fail(B200LP_E_INVALID, "slots is NULL but the structure has input planes");
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->pdhg) return pdhg_step_host(h, days, slots, node_flow, col_flow, objective);
    const size_t S = (size_t)h->S;
    if (h->ds.n_slots > 0) {
        const size_t bytes = sizeof(double) * S * h->ds.n_slots;
        CUDA_TRY(cudaMemcpyAsync(h->d_slots, slots, bytes, cudaMemcpyHostToDevice, h->stream));
        h->stats.h2d_bytes += (double)bytes;
    }
    h->h_fail[0] = 0; h->h_fail[1] = INT_MAX;
    CUDA_TRY(cudaMemcpyAsync(h->state.fail, h->h_fail, 2 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    int rc = launch_step(h, days, h->d_slots, node_flow ? h->d_node_flow : nullptr, col_flow ? h->d_col_flow : nullptr,
                         objective ? h->d_objective : nullptr);
    if (rc != 0) return rc;
    CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
    if (node_flow) {
        CUDA_TRY(cudaMemcpyAsync(node_flow, h->d_node_flow, sizeof(double) * S * h->ds.N, cudaMemcpyDeviceToHost, h->stream));
        h->stats.d2h_bytes += (double)(sizeof(double) * S * h->ds.N);
    }
    if (col_flow) {
        CUDA_TRY(cudaMemcpyAsync(col_flow, h->d_col_flow, sizeof(double) * S * h->ds.n, cudaMemcpyDeviceToHost, h->stream));
        h->stats.d2h_bytes += (double)(sizeof(double) * S * h->ds.n);
    }
    if (objective) {
        CUDA_TRY(cudaMemcpyAsync(objective, h->d_objective, sizeof(double) * S, cudaMemcpyDeviceToHost, h->stream));
        h->stats.d2h_bytes += (double)(sizeof(double) * S);
    }
    CUDA_TRY(cudaMemcpyAsync(h->h_fail, h->state.fail, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->stats.device_ms += ms;
    if (h->h_fail[0] > 0) {
        h->last_first_bad = h->h_fail[1] >> 4;
        h->last_code = h->h_fail[1] & 15;
        char buf[128];
        snprintf(buf, sizeof buf, "%d scenario(s) failed; first: scenario %d, status %d", h->h_fail[0], h->last_first_bad,
                 h->last_code);
        return fail(B200LP_E_SOLVE, buf);
    }
    h->last_first_bad = -1; h->last_code = 0;
    return B200LP_OK;
}

int b200lp_status(b200lp_handle *h, int32_t *first_bad_scenario, int32_t *code) {
    if (!h) return fail(B200LP_E_INVALID, "null handle");
    if (first_bad_scenario) *first_bad_scenario = h->last_first_bad;
    if (code) *code = h->last_code;
    return B200LP_OK;
}

int b200lp_get_basis(b200lp_handle *h, int32_t *row_stat, int32_t *col_stat) {
    if (!h || !row_stat || !col_stat) return fail(B200LP_E_INVALID, "null argument");
    if (h->pdhg && pdhg_is_rsx(h)) { CUDA_TRY(cudaSetDevice(h->device)); return rsx_get_basis(h, row_stat, col_stat); }
    if (h->pdhg) return fail(B200LP_E_UNSUPPORTED, "the PDHG iterations carry (x, y, omega) between timesteps, not a basis "
                                                   "(crossover disabled, or its state does not fit beside the PDHG state)");
    CUDA_TRY(cudaSetDevice(h->device));
    const int m = h->ds.m, n = h->ds.n, rows_cap = h->ds.rows_cap, nc = h->ds.nc;
    const size_t S = (size_t)h->S;
    std::vector<int> head(S * rows_cap), nb(S * nc), nbstat(S * nc);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(head.data(), h->state.head, sizeof(int) * head.size(), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(nb.data(), h->state.nb, sizeof(int) * nb.size(), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(nbstat.data(), h->state.nbstat, sizeof(int) * nbstat.size(), cudaMemcpyDeviceToHost));
    for (size_t s = 0; s < S; s++) {
        for (int i = 0; i < rows_cap; i++) {
            const int v = head[s * rows_cap + i];
            if (v < 0) continue;
            if (v < n) col_stat[s * n + v] = VS_BASIC; else row_stat[s * m + (v - n)] = VS_BASIC;
        }
        for (int j = 0; j < nc; j++) {
            const int v = nb[s * nc + j];
            if (v < 0) continue;
            if (v < n) col_stat[s * n + v] = nbstat[s * nc + j]; else row_stat[s * m + (v - n)] = nbstat[s * nc + j];
        }
    }
    return B200LP_OK;
}

int b200lp_get_volume(b200lp_handle *h, double *volume) {
    if (!h || !volume) return fail(B200LP_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(volume, h->state.vol, sizeof(double) * (size_t)h->ds.n_storage * h->S, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return B200LP_OK;
}

int b200lp_set_volume(b200lp_handle *h, const double *volume) {
    if (!h || !volume) return fail(B200LP_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(h->state.vol, volume, sizeof(double) * (size_t)h->ds.n_storage * h->S, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return B200LP_OK;
}

int b200lp_get_stats(b200lp_handle *h, b200lp_stats *out) {
    if (!h || !out) return fail(B200LP_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(h->h_counters, h->state.counters, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (h->pdhg && pdhg_is_rsx(h)) {
        if (h->pdhg->xover) h->stats.pdhg_iterations = (int64_t)h->h_counters[0];   // PDHG iterations + crossover pivots
        CUDA_TRY(cudaMemcpy(h->h_counters, pdhg_rsx_counters(h), 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        h->stats.pivots = (int64_t)h->h_counters[0];
        h->stats.refactors = (int64_t)h->h_counters[1];
    } else if (h->pdhg) {
        h->stats.pdhg_iterations = (int64_t)h->h_counters[0];
    } else {
        h->stats.pivots = (int64_t)h->h_counters[0];
        h->stats.refactors = (int64_t)h->h_counters[1];
    }
    *out = h->stats;
    return B200LP_OK;
}

int b200lp_set_option(b200lp_handle *h, const char *name, double value) {
    if (!h || !name) return fail(B200LP_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->pdhg) return pdhg_set_option(h, name, value);
    if (!strncmp(name, "pdhg_", 5)) return B200LP_OK;  // simplex engines ignore the PDHG options
    return fail(B200LP_E_INVALID, std::string("unknown option ") + name);
}

void *b200lp_stream(b200lp_handle *h) { return h ? (void *)h->stream : nullptr; }

int b200lp_sync(b200lp_handle *h) {
    if (!h) return fail(B200LP_E_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return B200LP_OK;
}

void *b200lp_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        g_last_error = "cudaHostAlloc failed";
        return nullptr;
    }
    return p;
}
void b200lp_host_free(void *p) { if (p) cudaFreeHost(p); }

void *b200lp_device_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
        g_last_error = "cudaMalloc failed";
        return nullptr;
    }
    return p;
}
void b200lp_device_free(void *p) { if (p) cudaFree(p); }

int b200lp_memcpy_h2d(b200lp_handle *h, void *d_dst, const void *src, size_t bytes) {
    if (!h) return fail(B200LP_E_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return B200LP_OK;
}

int b200lp_memcpy_d2h(b200lp_handle *h, void *dst, const void *d_src, size_t bytes) {
    if (!h) return fail(B200LP_E_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return B200LP_OK;
}

// ---------------------------------------------------------------------------------------------
// device-resident run
// ---------------------------------------------------------------------------------------------
}  // extern "C" (reopened below)

static void free_program(b200lp_handle *h) {
    plane_free(h);
    for (void *p : h->prog_owned) cudaFree(p);
    h->prog_owned.clear(); h->rec_ptr.clear(); h->rec_kind.clear();
    h->has_program = false;
    h->jspec.valid = false; h->jit_kernel = nullptr; h->jit_stale = false; h->stats.run_kernel_jit = 0;
}

template <typename T>
static cudaError_t pupload(b200lp_handle *h, const T *src, size_t count, const T **out) {
    T *p = nullptr;
    const size_t bytes = sizeof(T) * (count ? count : 1);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return e;
    h->prog_owned.push_back(p);
    e = cudaMemset(p, 0, bytes);
    if (e != cudaSuccess) return e;
    if (count && src) e = cudaMemcpy(p, src, sizeof(T) * count, cudaMemcpyHostToDevice);
    *out = p;
    return e;
}

// Ops of a program with every operand resolved to an index into V (pure host code, shared by b200lp_load_program and the
// specialised-kernel generator): table ops get one descriptor each, everything else is pre-decoded into `code` as
// kind, dst, nargs, args...
static bool decode_program(const b200lp_program *p, int n_slots, int n_consts, int nstor, int vol_base, int pc_base, int m, int n,
                           int S, const std::vector<int> &st_maxv_idx, DecodedProgram &dec, std::string &err) {
    auto bad = [&](const char *msg) { err = msg; return false; };
    // validate ops; table ops get one descriptor each, everything else is pre-decoded into `code`
    // with every operand resolved to an index into V
    const int nargs = p->n_ops ? p->op_arg_ptr[p->n_ops] : 0;
    auto ref_ok = [&](int ref) { return ref >= 0 ? ref < n_slots : (ref != B200LP_REF_STATE && (-ref - 1) < n_consts); };
    auto vidx = [&](int ref) { return ref >= 0 ? ref : n_slots + (-ref - 1); };
    std::vector<int> &tab_dst = dec.tab_dst, &tab_table = dec.tab_table, &tab_offset = dec.tab_offset, &code = dec.code;
    for (int k = 0; k < p->n_ops; k++) {
        if (p->op_dst[k] < 0 || p->op_dst[k] >= n_slots) return bad("op destination slot out of range");
        const int *a = p->op_args + p->op_arg_ptr[k];
        const int na = p->op_arg_ptr[k + 1] - p->op_arg_ptr[k];
        const int kind = p->op_kind[k];
        if (kind == B200LP_OP_TABLE) {
            if (na < 2) return bad("table op needs 2 arguments");
            const int tb = a[0];
            if (tb < 0 || tb >= p->n_tables) return bad("table index out of range");
            const int ax = p->table_axis[tb];
            if (ax >= p->n_axes) return bad("table axis out of range");
            if (ax == B200LP_COL_SCENARIO && p->table_cols[tb] < S) return bad("per-scenario table has too few columns");
            tab_dst.push_back(p->op_dst[k]); tab_table.push_back(tb); tab_offset.push_back(a[1]);
            continue;
        }
        std::vector<int> out;
        auto refs = [&](int from, int count) -> bool {
            if (from + count > na) return false;
            for (int i = 0; i < count; i++) {
                if (!ref_ok(a[from + i])) return false;
                out.push_back(vidx(a[from + i]));
            }
            return true;
        };
        bool ok = true;
        switch (kind) {
        case B200LP_OP_AGG:
            ok = na >= 2 && a[0] >= 0 && a[0] <= B200LP_AGG_MEAN && a[1] >= 1;
            if (ok) { out.push_back(a[0]); out.push_back(a[1]); ok = refs(2, a[1]); }
            break;
        case B200LP_OP_CONTROL_CURVE:
            ok = na >= 2 && a[0] >= 0 && a[0] < nstor && a[1] >= 0;
            if (ok) { out.push_back(a[0]); out.push_back(a[1]); ok = refs(2, 2 * a[1] + 1); }
            break;
        case B200LP_OP_CC_INDEX:
            ok = na >= 2 && a[0] >= 0 && a[0] < nstor && a[1] >= 0;
            if (ok) { out.push_back(a[0]); out.push_back(a[1]); ok = refs(2, a[1]); }
            break;
        case B200LP_OP_CC_INTERP:
            ok = na >= 2 && a[0] >= 0 && a[0] < nstor && a[1] >= 0;
            if (ok) { out.push_back(a[0]); out.push_back(a[1]); ok = refs(2, 2 * a[1] + 2); }
            break;
        case B200LP_OP_CC_PIECEWISE:
            ok = na >= 2 && a[0] >= 0 && a[0] < nstor && a[1] >= 0;
            if (ok) { out.push_back(a[0]); out.push_back(a[1]); ok = refs(2, 3 * a[1] + 4); }
            break;
        case B200LP_OP_INDEXED:
            ok = na >= 2 && a[1] >= 1 && ref_ok(a[0]);
            if (ok) { out.push_back(vidx(a[0])); out.push_back(a[1]); ok = refs(2, a[1]); }
            break;
        case B200LP_OP_NEG: ok = refs(0, 1); break;
        case B200LP_OP_MAX0: case B200LP_OP_MIN0: case B200LP_OP_DIV: ok = refs(0, 2); break;
        case B200LP_OP_THRESHOLD:  // -> V index of X, predicate, V indices of threshold, value0, value1
            ok = na >= 6 && a[2] >= B200LP_PRED_LT && a[2] <= B200LP_PRED_GE &&
                 ((a[0] == 0 && ref_ok(a[1])) || (a[0] == 1 && a[1] >= 0 && a[1] < nstor));
            if (ok) { out.push_back(a[0] == 0 ? vidx(a[1]) : vol_base + a[1]); out.push_back(a[2]); ok = refs(3, 3); }
            break;
        case B200LP_OP_INTERP:  // -> V index of X, n, V indices of the knots' x, y, below, above
            ok = na >= 3 && a[2] >= 2 && ((a[0] == 0 && ref_ok(a[1])) || (a[0] == 1 && a[1] >= 0 && a[1] < nstor));
            if (ok) { out.push_back(a[0] == 0 ? vidx(a[1]) : vol_base + a[1]); out.push_back(a[2]); ok = refs(3, 2 * a[2] + 2); }
            break;
        case B200LP_OP_STATE:  // -> V index of the volume / proportional volume, clamp flag
            ok = na >= 2 && a[0] >= 1 && a[0] <= 3 && a[1] >= 0 && a[1] < nstor;
            if (ok) { out.push_back(a[0] == 1 ? vol_base + a[1] : pc_base + a[1]); out.push_back(a[0] == 3 ? 1 : 0); }
            break;
        case B200LP_OP_LAG:  // -> recorder, V index of the initial value, steps
            ok = na >= 3 && a[0] >= 0 && a[0] < p->n_recorders && ref_ok(a[1]) && a[2] >= 1 && !(p->flags & B200LP_PROG_AGG_ONLY) &&
                 p->rec_kind[a[0]] >= B200LP_REC_NODE_FLOW && p->rec_kind[a[0]] <= B200LP_REC_STORAGE_PC;   // the [T, S] kinds
            if (ok) { out.push_back(a[0]); out.push_back(vidx(a[1])); out.push_back(a[2]); dec.has_lag = true; }
            break;
        case B200LP_OP_STORAGE_RESET:  // -> V indices: volume, pc, max_volume, flag, initial volume, initial pc
            ok = na >= 4 && a[0] >= 0 && a[0] < nstor;
            if (ok) {
                out.push_back(vol_base + a[0]); out.push_back(pc_base + a[0]); out.push_back(st_maxv_idx[a[0]]);
                ok = refs(1, 3);
            }
            break;
        default: ok = false;
        }
        if (!ok) return bad("malformed op");
        code.push_back(kind); code.push_back(p->op_dst[k]); code.push_back((int)out.size());
        code.insert(code.end(), out.begin(), out.end());
    }
    (void)nargs;
    std::vector<int> &rec_idx = dec.rec_idx, &rec_src = dec.rec_src;
    rec_idx.assign(p->n_recorders ? p->n_recorders : 1, 0); rec_src.assign(p->n_recorders ? p->n_recorders : 1, 0);
    for (int r = 0; r < p->n_recorders; r++) {
        const int kind = p->rec_kind[r];
        const bool needs_ref = kind == B200LP_REC_NODE_DEFICIT || kind == B200LP_REC_NODE_SUPPLIED || kind == B200LP_REC_NODE_CURTAIL ||
                               kind == B200LP_REC_TOTAL_DEFICIT || kind == B200LP_REC_DEFICIT_FREQ;
        if (needs_ref) {
            if (!ref_ok(p->rec_ref[r])) return bad("recorder max_flow ref out of range");
            rec_idx[r] = vidx(p->rec_ref[r]);
        }
        const bool storage_kind = kind == B200LP_REC_STORAGE_VOLUME || kind == B200LP_REC_STORAGE_PC || kind == B200LP_REC_MIN_VOLUME;
        if (storage_kind && (p->rec_src[r] < 0 || p->rec_src[r] >= nstor)) return bad("recorder storage index out of range");
        rec_src[r] = storage_kind ? p->rec_src[r] : 0;  // the kernel reads V[vol_base + src] unconditionally
        if (p->rec_row[r] >= m || p->rec_row[r] < -1) return bad("recorder row out of range");
        for (int e = p->rec_indptr[r]; e < p->rec_indptr[r + 1]; e++)
            if (p->rec_col[e] < 0 || p->rec_col[e] >= n) return bad("recorder column out of range");
    }
    return true;
}

static bool rec_is_array(int kind) { return kind >= B200LP_REC_NODE_FLOW && kind <= B200LP_REC_STORAGE_PC; }

#include "jit_run.inc"

// Blocks of 128 threads per SM the specialised kernel is compiled for (its register budget): the most that still puts the
// whole batch into one wave keeps the per-timestep chain shortest; large batches take the highest occupancy.
static int jit_pick_blocks(int grid) {
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    if (grid <= 2 * sms) return 2;
    if (grid <= 3 * sms) return 3;
    return 4;
}

// everything the generator needs from (resolved structure, program)
static void fill_spec(jit::Spec &sp, int G, int NC, const HostStructure &hs, const double *consts, const int *st_kind,
                      const b200lp_program *p, const DecodedProgram &dec) {
    sp = jit::Spec{};
    sp.G = G; sp.NC = NC; sp.hs = hs;
    sp.consts.assign(consts, consts + hs.n_consts);
    sp.st_kind.assign(st_kind, st_kind + hs.n_storage);
    sp.dec = dec;
    const int nr = p->n_recorders, nst = hs.n_storage;
    sp.n_recorders = nr;
    sp.rec_kind.assign(p->rec_kind, p->rec_kind + nr);
    sp.rec_row.assign(p->rec_row, p->rec_row + nr);
    sp.rec_indptr.assign(p->rec_indptr, p->rec_indptr + nr + 1);
    const int ne = nr ? p->rec_indptr[nr] : 0;
    sp.rec_col.assign(p->rec_col, p->rec_col + ne);
    sp.rec_w.assign(p->rec_w, p->rec_w + ne);
    sp.rec_factor.assign(p->rec_factor, p->rec_factor + nr);
    sp.stflow_indptr.assign(p->stflow_indptr, p->stflow_indptr + nst + 1);
    const int nf = nst ? p->stflow_indptr[nst] : 0;
    sp.stflow_col.assign(p->stflow_col, p->stflow_col + nf);
    sp.stflow_w.assign(p->stflow_w, p->stflow_w + nf);
    sp.roll_len.assign(nst ? nst : 1, 0); sp.roll_off.assign(nst ? nst : 1, 0);
    int rows = 0;
    for (int k = 0; k < nst; k++) {
        sp.roll_len[k] = p->st_roll_len ? std::max(0, p->st_roll_len[k]) : 0;
        sp.roll_off[k] = rows;
        rows += sp.roll_len[k];
    }
    sp.uniform_days = p->ts_days[0];
    for (int t = 1; t < p->n_steps; t++)
        if (p->ts_days[t] != sp.uniform_days) { sp.uniform_days = 0.0; break; }
    if (!(sp.uniform_days > 0.0)) sp.uniform_days = 0.0;
    sp.valid = true;
}

// (re)build the specialised kernel of the loaded program; leaves h->jit_kernel null when the model is not covered
static void jit_prepare(b200lp_handle *h) {
    h->jit_kernel = nullptr;
    h->stats.run_kernel_jit = 0;
    if (!h->jspec.valid) return;
    if (const char *v = getenv("B200LP_JIT")) if (atoi(v) == 0) return;
    std::string why;
    if (!jit::eligible(h->jspec, why)) { h->jit_note = why; return; }
    const int groups_per_block = h->block / h->jspec.G;
    const int grid = (h->S + groups_per_block - 1) / groups_per_block;
    int minb = jit_pick_blocks(grid);
    if (const char *v = getenv("B200LP_JIT_MINB")) { const int f = atoi(v); if (f >= 1 && f <= 8) minb = f; }
    h->jspec.minb = minb;
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device) != cudaSuccess || sms <= 0) sms = 148;
        const long long warps = ((long long)h->S * h->jspec.G + 31) / 32;
        h->jspec.defer_recorders = warps <= 4LL * sms;   // at most one warp per scheduler
        if (h->jspec.dec.has_lag) h->jspec.defer_recorders = false;   // a lagged sample must be in memory before the next timestep reads it
        if (const char *v = getenv("B200LP_JIT_DEFER")) h->jspec.defer_recorders = atoi(v) != 0;
    }
    const std::string src = jit::generate(h->jspec);
    if (const char *path = getenv("B200LP_JIT_DUMP")) { FILE *f = fopen(path, "w"); if (f) { fputs(src.c_str(), f); fclose(f); } }
    jit::Loaded *L = jit::get_kernel(src);
    if (!L) { h->jit_note = "run-time compilation failed"; return; }
    h->jit_kernel = L->kernel;
    h->stats.run_kernel_jit = 1;
    h->jit_note.clear();
}

extern "C" {

int b200lp_jit_probe(const b200lp_structure *s, const b200lp_program *p, int32_t n_scenarios, int32_t compile, char *log, size_t cap) {
    if (!s || !p) return fail(B200LP_E_INVALID, "null argument");
    if (log && cap) log[0] = 0;
    int variant = -1;
    for (int v = 0; v < kNumVariants; v++)
        if (s->n_rows <= kVariants[v].G * kVariants[v].RPL && s->n_cols <= kVariants[v].NC) { variant = v; break; }
    if (variant < 0 || kVariants[variant].RPL != 1) return fail(B200LP_E_UNSUPPORTED, "no one-row-per-lane register variant fits this LP");
    const Variant &V = kVariants[variant];
    HostStructure hs;
    std::string err;
    if (!resolve_structure(s, V.G * V.RPL, hs, err)) return fail(B200LP_E_INVALID, err);
    DecodedProgram dec;
    if (!decode_program(p, hs.n_slots, hs.n_consts, hs.n_storage, hs.vol_base, hs.pc_base, hs.m, hs.n, n_scenarios, hs.st_maxv, dec, err))
        return fail(B200LP_E_INVALID, err);
    jit::Spec sp;
    fill_spec(sp, V.G, V.NC, hs, s->consts, s->st_kind, p, dec);
    if (!jit::eligible(sp, err)) return fail(B200LP_E_UNSUPPORTED, "not specialised: " + err);
    const std::string src = jit::generate(sp);
    if (const char *path = getenv("B200LP_JIT_DUMP")) { FILE *f = fopen(path, "w"); if (f) { fputs(src.c_str(), f); fclose(f); } }
    if (!compile) return B200LP_OK;
    std::vector<char> cubin;
    std::string clog;
    const bool ok = jit::compile(src, cubin, clog);
    if (log && cap) { strncpy(log, clog.c_str(), cap - 1); log[cap - 1] = 0; }
    if (!ok) return fail(B200LP_E_CUDA, "NVRTC: " + clog.substr(0, 400));
    if (const char *path = getenv("B200LP_JIT_DUMP")) {  // <path>.cubin for cuobjdump -sass
        FILE *f = fopen((std::string(path) + ".cubin").c_str(), "wb");
        if (f) { fwrite(cubin.data(), 1, cubin.size(), f); fclose(f); }
    }
    return B200LP_OK;
}

int b200lp_load_program(b200lp_handle *h, const b200lp_program *p) {
    if (!h || !p) return fail(B200LP_E_INVALID, "null argument");
    if (p->n_steps <= 0) return fail(B200LP_E_INVALID, "program has no timesteps");
    if (h->use_cta && !h->pdhg)
        return fail(B200LP_E_UNSUPPORTED, "the device-resident run needs a register-resident kernel variant; "
                                          "this LP uses the CTA-per-scenario kernel (per-timestep path only)");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    free_program(h);
    const size_t S = (size_t)h->S;
    const int nst = h->ds.n_storage;
    DevProgram &dp = h->dp;
    dp = DevProgram{};
    dp.n_steps = p->n_steps; dp.n_tables = p->n_tables; dp.n_axes = p->n_axes;
    dp.n_recorders = p->n_recorders;
    DecodedProgram dec;
    {
        std::string err;
        if (!decode_program(p, h->ds.n_slots, h->ds.n_consts, h->ds.n_storage, h->ds.vol_base, h->ds.pc_base, h->ds.m, h->ds.n,
                            h->S, h->h_st_maxv_idx, dec, err))
            return fail(B200LP_E_INVALID, err);
    }
    const int n_slots = h->ds.n_slots;
    std::vector<int> &tab_dst = dec.tab_dst, &tab_table = dec.tab_table, &tab_offset = dec.tab_offset, &code = dec.code;
    std::vector<int> &rec_idx = dec.rec_idx, &rec_src = dec.rec_src;
    dp.n_table_ops = (int)tab_dst.size();
    dp.n_code = (int)code.size();
    // ---- lane-parallel schedule of the non-table ops (falls back to the interpreter when an op does not fit)
    std::vector<int> fop;
    int n_levels = 0;
    {
        const int KB = h->ds.k_base;
        const int G = kVariants[h->variant].G;
        std::vector<int> slot_level(n_slots, 0);
        struct FastOp { int level; int w[FOP_WORDS]; };
        std::vector<FastOp> ops;
        bool ok = true;
        size_t pos = 0;
        while (ok && pos < code.size()) {
            const int kind = code[pos], dst = code[pos + 1], na = code[pos + 2];
            const int *a = code.data() + pos + 3;
            pos += 3 + na;
            FastOp f{}; f.w[1] = dst;
            std::vector<int> deps;
            auto pad = [&](int from, int neutral) { for (int i = from; i < 6; i++) f.w[2 + i] = KB + neutral; };
            switch (kind) {
            case B200LP_OP_AGG: {
                const int func = a[0], cnt = a[1];
                if (cnt > 6 || func == B200LP_AGG_MEAN) { ok = false; break; }
                f.w[0] = func == B200LP_AGG_PRODUCT ? FOP_PROD : func == B200LP_AGG_SUM ? FOP_SUM : func == B200LP_AGG_MIN ? FOP_MIN : FOP_MAX;
                for (int i = 0; i < cnt; i++) { f.w[2 + i] = a[2 + i]; deps.push_back(a[2 + i]); }
                pad(cnt, func == B200LP_AGG_PRODUCT ? K_ONE : func == B200LP_AGG_SUM ? K_ZERO : func == B200LP_AGG_MIN ? K_PINF : K_NINF);
            } break;
            case B200LP_OP_CONTROL_CURVE: case B200LP_OP_CC_INDEX: {
                const int cnt = a[1];
                if (cnt > 2) { ok = false; break; }
                f.w[0] = FOP_CC;
                f.w[2] = h->ds.pc_base + a[0];
                f.w[3] = cnt >= 1 ? a[2] : KB + K_NINF;   // no curve: always the first value
                f.w[4] = cnt >= 2 ? a[3] : KB + K_PINF;
                for (int i = 0; i < cnt; i++) deps.push_back(a[2 + i]);
                if (kind == B200LP_OP_CONTROL_CURVE) {
                    const int *v = a + 2 + cnt;             // cnt + 1 values
                    f.w[5] = v[0]; f.w[6] = cnt >= 1 ? v[1] : v[0]; f.w[7] = cnt >= 2 ? v[2] : (cnt >= 1 ? v[1] : v[0]);
                    for (int i = 0; i <= cnt; i++) deps.push_back(v[i]);
                } else {
                    f.w[5] = KB + K_ZERO; f.w[6] = cnt >= 1 ? KB + K_ONE : KB + K_ZERO; f.w[7] = cnt >= 2 ? KB + K_TWO : f.w[6];
                }
            } break;
            case B200LP_OP_INDEXED: {
                const int cnt = a[1];
                if (cnt > 4) { ok = false; break; }
                f.w[0] = FOP_INDEXED; f.w[2] = a[0]; deps.push_back(a[0]);
                for (int i = 0; i < 4; i++) { f.w[3 + i] = a[2 + (i < cnt ? i : cnt - 1)]; }
                for (int i = 0; i < cnt; i++) deps.push_back(a[2 + i]);
                f.w[7] = KB + K_ZERO;
            } break;
            case B200LP_OP_NEG: f.w[0] = FOP_NEG; f.w[2] = a[0]; deps.push_back(a[0]); pad(1, K_ZERO); break;
            case B200LP_OP_MAX0: case B200LP_OP_MIN0: case B200LP_OP_DIV:
                f.w[0] = kind == B200LP_OP_MAX0 ? FOP_MAX0 : kind == B200LP_OP_MIN0 ? FOP_MIN0 : FOP_DIV;
                f.w[2] = a[0]; f.w[3] = a[1]; deps.push_back(a[0]); deps.push_back(a[1]); pad(2, K_ONE);
                break;
            default: ok = false;
            }
            if (!ok) break;
            int lev = 0;
            for (int dref : deps) if (dref < n_slots) lev = std::max(lev, slot_level[dref]);
            // an op that reads a storage's pc / volume does so at level >= 0 too; ops writing dst raise its level
            f.level = lev;
            slot_level[dst] = lev + 1;
            ops.push_back(f);
        }
        if (ok && !ops.empty()) {
            int maxlev = 0;
            for (auto &f : ops) maxlev = std::max(maxlev, f.level);
            n_levels = maxlev + 1;
            std::vector<int> count(n_levels, 0);
            if (n_levels > FOP_MAXLEV) ok = false;
            for (auto &f : ops) if (++count[f.level] > G) ok = false;
            if (ok) {
                fop.assign((size_t)n_levels * 32 * FOP_WORDS, 0);
                for (int lev = 0; lev < n_levels; lev++)
                    for (int lane = 0; lane < 32; lane++) {
                        int *w = fop.data() + ((size_t)lev * 32 + lane) * FOP_WORDS;
                        w[0] = FOP_NONE; w[1] = KB + K_SCRATCH;
                        for (int i = 2; i < FOP_WORDS; i++) w[i] = KB + K_ONE;
                    }
                std::fill(count.begin(), count.end(), 0);
                for (auto &f : ops) {
                    int *w = fop.data() + ((size_t)f.level * 32 + count[f.level]++) * FOP_WORDS;
                    for (int i = 0; i < FOP_WORDS; i++) w[i] = f.w[i];
                }
            }
        }
        if (!ok) { n_levels = 0; fop.clear(); }
        if (getenv("B200LP_NO_FAST_OPS")) { n_levels = 0; fop.clear(); }
    }
    dp.n_fop_levels = n_levels;
    long long ndata = 0;
    for (int t = 0; t < p->n_tables; t++) {
        const long long end = (long long)p->table_offset[t] + (long long)p->table_rows[t] * p->table_cols[t];
        if (end > ndata) ndata = end;
    }
    // scenario indices must address existing columns
    for (int t = 0; t < p->n_tables; t++) {
        const int ax = p->table_axis[t];
        if (ax >= p->n_axes || ax < B200LP_COL_SCENARIO) return fail(B200LP_E_INVALID, "table axis out of range");
        if (p->table_rows[t] < 1 || p->table_cols[t] < 1) return fail(B200LP_E_INVALID, "empty table");
        if (ax >= 0)
            for (size_t s = 0; s < S; s++)
                if (p->scen_index[(size_t)ax * S + s] < 0 || p->scen_index[(size_t)ax * S + s] >= p->table_cols[t])
                    return fail(B200LP_E_INVALID, "scenario index exceeds table columns");
    }
#define PUP(field, src, count)                                                          \
    do {                                                                                \
        cudaError_t _e = pupload(h, src, (size_t)(count), &dp.field);                   \
        if (_e != cudaSuccess) {                                                        \
            free_program(h);                                                            \
            return fail(B200LP_E_CUDA, std::string("program upload " #field ": ") + cudaGetErrorString(_e)); \
        }                                                                               \
    } while (0)
    PUP(ts_days, p->ts_days, p->n_steps);
    PUP(tab_dst, tab_dst.data(), tab_dst.size());
    PUP(tab_table, tab_table.data(), tab_table.size());
    PUP(tab_offset, tab_offset.data(), tab_offset.size());
    PUP(code, code.data(), code.size());
    PUP(fop, fop.data(), fop.size());
    PUP(table_rows, p->table_rows, p->n_tables);
    PUP(table_cols, p->table_cols, p->n_tables);
    PUP(table_axis, p->table_axis, p->n_tables);
    PUP(table_offset, reinterpret_cast<const long long *>(p->table_offset), p->n_tables);
    PUP(table_data, p->table_data, ndata);
    PUP(scen_index, p->scen_index, (size_t)p->n_axes * S);
    PUP(stflow_indptr, p->stflow_indptr, nst + 1);
    PUP(stflow_col, p->stflow_col, nst ? p->stflow_indptr[nst] : 0);
    PUP(stflow_w, p->stflow_w, nst ? p->stflow_indptr[nst] : 0);
    PUP(rec_kind, p->rec_kind, p->n_recorders);
    PUP(rec_src, rec_src.data(), p->n_recorders);
    PUP(rec_idx, rec_idx.data(), p->n_recorders);
    PUP(rec_row, p->rec_row, p->n_recorders);
    PUP(rec_indptr, p->rec_indptr, p->n_recorders + 1);
    PUP(rec_col, p->rec_col, p->n_recorders ? p->rec_indptr[p->n_recorders] : 0);
    PUP(rec_factor, p->rec_factor, p->n_recorders);
    PUP(rec_w, p->rec_w, p->n_recorders ? p->rec_indptr[p->n_recorders] : 0);
#undef PUP
    h->stats.h2d_bytes += (double)ndata * 8.0;
    auto palloc = [&](size_t bytes, void **out) -> cudaError_t {
        cudaError_t e = cudaMalloc(out, bytes ? bytes : 8);
        if (e == cudaSuccess) h->prog_owned.push_back(*out);
        return e;
    };
    h->rec_kind.assign(p->rec_kind, p->rec_kind + p->n_recorders);
    h->rec_ptr.resize(p->n_recorders);
    h->agg_only = (p->flags & B200LP_PROG_AGG_ONLY) != 0;
    for (int r = 0; r < p->n_recorders; r++) {
        const size_t count = rec_is_array(p->rec_kind[r]) ? (size_t)p->n_steps * S : S;
        void *q = nullptr;
        h->rec_ptr[r] = nullptr;
        if (rec_is_array(p->rec_kind[r]) && h->agg_only) continue;
        if (palloc(sizeof(double) * count, &q) != cudaSuccess) { free_program(h); return fail(B200LP_E_CUDA, "recorder output allocation failed"); }
        h->rec_ptr[r] = (double *)q;
    }
    void *q = nullptr;
    if (palloc(sizeof(double) * 3 * S * (p->n_recorders ? p->n_recorders : 1), &q) != cudaSuccess) { free_program(h); return fail(B200LP_E_CUDA, "alloc"); }
    dp.rec_agg = (double *)q;
    if (palloc(sizeof(double *) * (p->n_recorders + 1), &q) != cudaSuccess) { free_program(h); return fail(B200LP_E_CUDA, "alloc"); }
    if (p->n_recorders) cudaMemcpy(q, h->rec_ptr.data(), sizeof(double *) * p->n_recorders, cudaMemcpyHostToDevice);
    dp.rec_out = (double *const *)q;
    if (palloc(sizeof(double) * S * (nst ? nst : 1), &q) != cudaSuccess) { free_program(h); return fail(B200LP_E_CUDA, "alloc"); }
    dp.pc = (double *)q;
    if (palloc(sizeof(double) * S * h->ds.n, &q) != cudaSuccess) { free_program(h); return fail(B200LP_E_CUDA, "alloc"); }
    dp.last_x = (double *)q;
    if (palloc(sizeof(int) * S, &q) != cudaSuccess) { free_program(h); return fail(B200LP_E_CUDA, "alloc"); }
    dp.fail_step = (int *)q;
    // rolling virtual storages: one ring buffer [len][S] each, rows stacked in one allocation
    h->roll_len.assign(nst ? nst : 1, 0); h->roll_off.assign(nst ? nst : 1, 0); h->roll_init.assign(nst ? nst : 1, 0.0);
    dp.roll_len = nullptr; dp.roll_off = nullptr; dp.roll_mem = nullptr;
    {
        int rows = 0;
        for (int k = 0; k < nst; k++) {
            h->roll_len[k] = p->st_roll_len ? std::max(0, p->st_roll_len[k]) : 0;
            h->roll_init[k] = (p->st_roll_init && h->roll_len[k] > 0) ? p->st_roll_init[k] : 0.0;
            h->roll_off[k] = rows;
            rows += h->roll_len[k];
        }
        if (rows > 0) {
            if (palloc(sizeof(double) * (size_t)rows * S, &q) != cudaSuccess) { free_program(h); return fail(B200LP_E_CUDA, "alloc"); }
            dp.roll_mem = (double *)q;
            if (palloc(sizeof(int) * nst, &q) != cudaSuccess) { free_program(h); return fail(B200LP_E_CUDA, "alloc"); }
            cudaMemcpy(q, h->roll_len.data(), sizeof(int) * nst, cudaMemcpyHostToDevice);
            dp.roll_len = (const int *)q;
            if (palloc(sizeof(int) * nst, &q) != cudaSuccess) { free_program(h); return fail(B200LP_E_CUDA, "alloc"); }
            cudaMemcpy(q, h->roll_off.data(), sizeof(int) * nst, cudaMemcpyHostToDevice);
            dp.roll_off = (const int *)q;
        }
    }
    h->table_rows.assign(p->table_rows, p->table_rows + p->n_tables);
    h->table_cols.assign(p->table_cols, p->table_cols + p->n_tables);
    h->table_offset.assign(p->table_offset, p->table_offset + p->n_tables);
    h->init_vol.assign(p->initial_volume, p->initial_volume + (size_t)nst * S);
    h->init_pc.assign(p->initial_pc, p->initial_pc + (size_t)nst * S);
    h->prog_steps = p->n_steps;
    h->has_program = true;
    if (!h->pdhg && !h->use_cta && kVariants[h->variant].RPL == 1) {
        fill_spec(h->jspec, kVariants[h->variant].G, kVariants[h->variant].NC, h->hs, h->consts_host.data(), h->st_kind_host.data(), p, dec);
        jit_prepare(h);
    }
    if (h->pdhg) {  // large-LP path: the program runs as plane kernels around the batched solve (plane_program.inc)
        const int rc = plane_attach(h, p);
        if (rc != 0) { free_program(h); return rc; }
    }
    return b200lp_program_reset(h);
}

int b200lp_program_reset(b200lp_handle *h) {
    if (!h || !h->has_program) return fail(B200LP_E_INVALID, "no program loaded");
    int rc = b200lp_reset(h, h->ds.n_storage ? h->init_vol.data() : nullptr);
    if (rc != 0) return rc;
    const size_t S = (size_t)h->S;
    if (h->ds.n_storage)
        CUDA_TRY(cudaMemcpyAsync(h->dp.pc, h->init_pc.data(), sizeof(double) * S * h->ds.n_storage, cudaMemcpyHostToDevice, h->stream));
    if (!h->rec_ptr.empty()) {
        const size_t count = 3 * S * h->rec_ptr.size();
        agg_reset_kernel<<<(unsigned)((count + 255) / 256), 256, 0, h->stream>>>(h->dp.rec_agg, count, S);
        h->stats.kernel_launches++;
    }
    for (size_t r = 0; r < h->rec_ptr.size(); r++) {
        const size_t count = rec_is_array(h->rec_kind[r]) ? (size_t)h->prog_steps * S : S;
        if (!h->rec_ptr[r]) continue;
        if (h->rec_kind[r] == B200LP_REC_MIN_VOLUME) {
            fill_kernel<<<(unsigned)((count + 255) / 256), 256, 0, h->stream>>>(h->rec_ptr[r], count, INFINITY);
            h->stats.kernel_launches++;
        } else {
            CUDA_TRY(cudaMemsetAsync(h->rec_ptr[r], 0, sizeof(double) * count, h->stream));
        }
    }
    if (h->dp.roll_mem)
        for (int k = 0; k < h->ds.n_storage; k++)
            if (h->roll_len[k] > 0) {
                const size_t count = (size_t)h->roll_len[k] * S;
                fill_kernel<<<(unsigned)((count + 255) / 256), 256, 0, h->stream>>>(h->dp.roll_mem + (size_t)h->roll_off[k] * S, count,
                                                                                 h->roll_init[k]);
                h->stats.kernel_launches++;
            }
    CUDA_TRY(cudaMemsetAsync(h->dp.last_x, 0, sizeof(double) * S * h->ds.n, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->dp.fail_step, 0xff, sizeof(int) * S, h->stream));
    if (h->plane) CUDA_TRY(cudaMemsetAsync(h->plane->Q.dead, 0, sizeof(int) * S, h->stream));
    h->h_fail[0] = 0; h->h_fail[1] = INT_MAX;
    CUDA_TRY(cudaMemcpyAsync(h->state.fail, h->h_fail, 2 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return B200LP_OK;
}

int b200lp_run(b200lp_handle *h, int32_t t0, int32_t n) {
    if (!h || !h->has_program) return fail(B200LP_E_INVALID, "no program loaded");
    if (t0 < 0 || n <= 0 || t0 + n > h->prog_steps) return fail(B200LP_E_INVALID, "timestep range outside the program");
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->plane) return plane_run(h, t0, n);
    const Variant &V = kVariants[h->variant];
    if (h->jit_stale) { h->jit_stale = false; jit_prepare(h); }
    if (h->jit_kernel) {
        const int groups_per_block = h->block / V.G;
        const int grid = (h->S + groups_per_block - 1) / groups_per_block;
        const size_t smem = ((V.smem_per_group(h->ds.n_val) + 15) & ~(size_t)15) * groups_per_block;
        if (smem > 48 * 1024) ensure_smem((const void *)h->jit_kernel, smem);
        int S = h->S, t0_ = t0, n_ = n;
        void *args[] = {&h->ds, &h->state, &h->dp, &S, &t0_, &n_};
        CUDA_TRY(cudaLaunchKernel((const void *)h->jit_kernel, dim3(grid), dim3(h->block), args, smem, h->stream));
    } else {
        V.run(h->ds, h->state, h->dp, h->S, t0, n, h->stream, h->block);
    }
    h->stats.kernel_launches++;
    h->stats.steps += n;
    h->stats.scenario_steps += (int64_t)n * h->S;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(B200LP_E_CUDA, std::string("run kernel launch: ") + cudaGetErrorString(e));
    return B200LP_OK;
}

static int run_pipelined_body(b200lp_handle *h, int32_t n_chunks, const double *const *table_src, double *const *rec_dst);

int b200lp_run_pipelined(b200lp_handle *h, int32_t n_chunks, const double *const *table_src, double *const *rec_dst) {
    if (!h || !h->has_program) return fail(B200LP_E_INVALID, "no program loaded");
    CUDA_TRY(cudaSetDevice(h->device));
    const int rc = run_pipelined_body(h, n_chunks, table_src, rec_dst);
    if (rc != 0) {
        // no copy into or out of the caller's buffers may still be in flight when the call returns
        const std::string msg = g_last_error;
        if (h->s_in) cudaStreamSynchronize(h->s_in);
        cudaStreamSynchronize(h->stream);
        if (h->s_out) cudaStreamSynchronize(h->s_out);
        cudaGetLastError();
        g_last_error = msg;
    }
    return rc;
}

static int run_pipelined_body(b200lp_handle *h, int32_t n_chunks, const double *const *table_src, double *const *rec_dst) {
    const int T = h->prog_steps;
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > T) n_chunks = T;
    if (!h->s_in) CUDA_TRY(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    if (!h->s_out) CUDA_TRY(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    while ((int)h->pipe_events.size() < 2 * n_chunks + 1) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->pipe_events.push_back(e);
    }
    int rc = b200lp_program_reset(h);  // Model.reset: initial volumes, slack bases, cleared recorders (synchronises)
    if (rc != 0) return rc;
    const size_t S = (size_t)h->S;
    double *tables = const_cast<double *>(h->dp.table_data);
    const int n_tables = h->dp.n_tables, n_rec = (int)h->rec_ptr.size();
    // tables that do not advance with time go up first, whole
    if (table_src)
        for (int k = 0; k < n_tables; k++)
            if (table_src[k] && h->table_rows[k] != T) {
                const size_t count = (size_t)h->table_rows[k] * h->table_cols[k];
                CUDA_TRY(cudaMemcpyAsync(tables + h->table_offset[k], table_src[k], sizeof(double) * count, cudaMemcpyHostToDevice, h->s_in));
                h->stats.h2d_bytes += (double)count * 8.0;
            }
    int rows_up = 0;  // rows [0, rows_up) of the time-indexed tables are on the device (or queued)
    for (int c = 0; c < n_chunks; c++) {
        const int t0 = (int)((long long)T * c / n_chunks), t1 = (int)((long long)T * (c + 1) / n_chunks);
        // the kernel reads one row ahead (prefetch of timestep t+1, clamped to the last row)
        const int need = std::min(T, t1 + 1);
        if (table_src && need > rows_up)
            for (int k = 0; k < n_tables; k++)
                if (table_src[k] && h->table_rows[k] == T) {
                    const size_t cols = (size_t)h->table_cols[k], off = (size_t)rows_up * cols, count = (size_t)(need - rows_up) * cols;
                    CUDA_TRY(cudaMemcpyAsync(tables + h->table_offset[k] + off, table_src[k] + off, sizeof(double) * count,
                                             cudaMemcpyHostToDevice, h->s_in));
                    h->stats.h2d_bytes += (double)count * 8.0;
                }
        rows_up = std::max(rows_up, need);
        cudaEvent_t ev_in = h->pipe_events[2 * c], ev_run = h->pipe_events[2 * c + 1];
        CUDA_TRY(cudaEventRecord(ev_in, h->s_in));
        CUDA_TRY(cudaStreamWaitEvent(h->stream, ev_in, 0));
        rc = b200lp_run(h, t0, t1 - t0);
        if (rc != 0) return rc;
        CUDA_TRY(cudaEventRecord(ev_run, h->stream));
        CUDA_TRY(cudaStreamWaitEvent(h->s_out, ev_run, 0));
        if (rec_dst)
            for (int r = 0; r < n_rec; r++) {
                if (!rec_dst[r] || !h->rec_ptr[r]) continue;
                if (rec_is_array(h->rec_kind[r])) {
                    const size_t off = (size_t)t0 * S, count = (size_t)(t1 - t0) * S;
                    CUDA_TRY(cudaMemcpyAsync(rec_dst[r] + off, h->rec_ptr[r] + off, sizeof(double) * count, cudaMemcpyDeviceToHost, h->s_out));
                    h->stats.d2h_bytes += (double)count * 8.0;
                } else if (c == n_chunks - 1) {
                    CUDA_TRY(cudaMemcpyAsync(rec_dst[r], h->rec_ptr[r], sizeof(double) * S, cudaMemcpyDeviceToHost, h->s_out));
                    h->stats.d2h_bytes += (double)S * 8.0;
                }
            }
    }
    CUDA_TRY(cudaStreamSynchronize(h->s_out));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->s_in));
    return B200LP_OK;
}

int b200lp_program_status(b200lp_handle *h, int32_t *n_failed, int32_t *scenario, int32_t *code, int32_t *timestep) {
    if (!h || !h->has_program) return fail(B200LP_E_INVALID, "no program loaded");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(h->h_fail, h->state.fail, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    int sc = -1, cd = 0, ts = -1;
    if (h->h_fail[0] > 0) {
        sc = h->h_fail[1] >> 4; cd = h->h_fail[1] & 15;
        CUDA_TRY(cudaMemcpy(&ts, h->dp.fail_step + sc, sizeof(int), cudaMemcpyDeviceToHost));
    }
    if (n_failed) *n_failed = h->h_fail[0];
    if (scenario) *scenario = sc;
    if (code) *code = cd;
    if (timestep) *timestep = ts;
    return B200LP_OK;
}

int b200lp_update_table(b200lp_handle *h, int32_t table, const double *data) {
    if (!h || !h->has_program || !data) return fail(B200LP_E_INVALID, "no program loaded / null data");
    if (table < 0 || table >= h->dp.n_tables) return fail(B200LP_E_INVALID, "table index out of range");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t count = (size_t)h->table_rows[table] * h->table_cols[table];
    CUDA_TRY(cudaMemcpyAsync(const_cast<double *>(h->dp.table_data) + h->table_offset[table], data, sizeof(double) * count,
                             cudaMemcpyHostToDevice, h->stream));
    h->stats.h2d_bytes += (double)count * 8.0;
    return B200LP_OK;
}

int b200lp_mark(b200lp_handle *h, int32_t which) {
    if (!h) return fail(B200LP_E_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaEventRecord(which ? h->ev1 : h->ev0, h->stream));
    return B200LP_OK;
}

int b200lp_elapsed_ms(b200lp_handle *h, double *ms) {
    if (!h || !ms) return fail(B200LP_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaEventSynchronize(h->ev1));
    float f = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&f, h->ev0, h->ev1));
    *ms = (double)f;
    return B200LP_OK;
}

int b200lp_fetch_recorder(b200lp_handle *h, int32_t rec, double *out) {
    if (!h || !h->has_program || !out) return fail(B200LP_E_INVALID, "no program loaded / null output");
    if (rec < 0 || rec >= (int)h->rec_ptr.size()) return fail(B200LP_E_INVALID, "recorder index out of range");
    CUDA_TRY(cudaSetDevice(h->device));
    if (!h->rec_ptr[rec]) return fail(B200LP_E_UNSUPPORTED, "the program was loaded with B200LP_PROG_AGG_ONLY: this array recorder keeps aggregates only");
    const size_t count = rec_is_array(h->rec_kind[rec]) ? (size_t)h->prog_steps * h->S : (size_t)h->S;
    CUDA_TRY(cudaMemcpyAsync(out, h->rec_ptr[rec], sizeof(double) * count, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->stats.d2h_bytes += (double)count * 8.0;
    return B200LP_OK;
}

int b200lp_fetch_recorder_strided(b200lp_handle *h, int32_t rec, double *out, int64_t ld) {
    if (!h || !h->has_program || !out) return fail(B200LP_E_INVALID, "no program loaded / null output");
    if (rec < 0 || rec >= (int)h->rec_ptr.size()) return fail(B200LP_E_INVALID, "recorder index out of range");
    if (ld < h->S) return fail(B200LP_E_INVALID, "leading dimension smaller than the scenario count");
    CUDA_TRY(cudaSetDevice(h->device));
    if (!h->rec_ptr[rec]) return fail(B200LP_E_UNSUPPORTED, "the program was loaded with B200LP_PROG_AGG_ONLY: this array recorder keeps aggregates only");
    const size_t rows = rec_is_array(h->rec_kind[rec]) ? (size_t)h->prog_steps : 1, S = (size_t)h->S;
    CUDA_TRY(cudaMemcpy2DAsync(out, sizeof(double) * (size_t)ld, h->rec_ptr[rec], sizeof(double) * S, sizeof(double) * S, rows,
                               cudaMemcpyDeviceToHost, h->stream));
    h->stats.d2h_bytes += (double)(rows * S) * 8.0;
    return B200LP_OK;
}

int b200lp_fetch_recorder_agg(b200lp_handle *h, int32_t rec, double *sum, double *mn, double *mx) {
    if (!h || !h->has_program) return fail(B200LP_E_INVALID, "no program loaded");
    if (rec < 0 || rec >= (int)h->rec_ptr.size()) return fail(B200LP_E_INVALID, "recorder index out of range");
    if (!rec_is_array(h->rec_kind[rec])) return fail(B200LP_E_INVALID, "not an array recorder: use b200lp_fetch_recorder");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t S = (size_t)h->S;
    double *dst[3] = {sum, mn, mx};
    for (int k = 0; k < 3; k++)
        if (dst[k]) {
            CUDA_TRY(cudaMemcpyAsync(dst[k], h->dp.rec_agg + ((size_t)rec * 3 + k) * S, sizeof(double) * S, cudaMemcpyDeviceToHost, h->stream));
            h->stats.d2h_bytes += (double)S * 8.0;
        }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return B200LP_OK;
}

void *b200lp_recorder_device_ptr(b200lp_handle *h, int32_t rec, int64_t *count) {
    if (!h || !h->has_program || rec < 0 || rec >= (int)h->rec_ptr.size()) return nullptr;
    if (count) *count = !h->rec_ptr[rec] ? 0 : (rec_is_array(h->rec_kind[rec]) ? (int64_t)h->prog_steps * h->S : (int64_t)h->S);
    return h->rec_ptr[rec];
}

int b200lp_fetch_state(b200lp_handle *h, double *col_flow, double *volume, double *pc) {
    if (!h) return fail(B200LP_E_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t S = (size_t)h->S;
    if (col_flow) {
        if (!h->has_program) return fail(B200LP_E_INVALID, "no program loaded");
        CUDA_TRY(cudaMemcpyAsync(col_flow, h->dp.last_x, sizeof(double) * S * h->ds.n, cudaMemcpyDeviceToHost, h->stream));
    }
    if (volume && h->ds.n_storage)
        CUDA_TRY(cudaMemcpyAsync(volume, h->state.vol, sizeof(double) * S * h->ds.n_storage, cudaMemcpyDeviceToHost, h->stream));
    if (pc && h->ds.n_storage) {
        if (!h->has_program) return fail(B200LP_E_INVALID, "no program loaded");
        CUDA_TRY(cudaMemcpyAsync(pc, h->dp.pc, sizeof(double) * S * h->ds.n_storage, cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return B200LP_OK;
}

#ifdef RSX_PHASE_TIMING
int b200lp_debug_rsx_phases(b200lp_handle *h, unsigned long long *out16) {  /* 18 values */
    if (!h || !h->pdhg || !pdhg_is_rsx(h)) return -1;
    return cudaMemcpy(out16, pdhg_rsx_counters(h), sizeof(unsigned long long) * 18, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -3;
}
#endif

#ifdef B200LP_PHASE_TIMING
int b200lp_debug_phase_cycles(unsigned long long *out16) {
    return cudaMemcpyFromSymbol(out16, b200lp::g_phase_cycles, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -3;
}
#endif

}  // extern "C"
