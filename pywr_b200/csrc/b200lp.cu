// b200lp.cu — C-ABI (include/b200lp.h) of the B200 scenario-batched LP engine: handle management,
// device memory, host<->device copies and kernel dispatch.  No torch, no Python, no CPU fallback.
#include <cuda_runtime.h>

#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200lp.h"
#include "lp_kernel.cuh"

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

typedef void (*launch_fn)(const DevStructure &, const DevState &, int S, double days, const double *slots,
                          double *node_flow, double *col_flow, double *objective, cudaStream_t stream, int block);

template <int G, int RPL, int NC>
static void launch_variant(const DevStructure &st, const DevState &state, int S, double days, const double *slots,
                           double *node_flow, double *col_flow, double *objective, cudaStream_t stream, int block) {
    const int groups_per_block = block / G;
    const int grid = (S + groups_per_block - 1) / groups_per_block;
    const size_t gbytes = (GroupSmem<G, RPL, NC>::bytes(st.n_slots) + 15) & ~(size_t)15;
    const size_t smem = gbytes * groups_per_block;
    static bool attr_set = false;
    if (!attr_set && smem > 48 * 1024) {
        cudaFuncSetAttribute(lp_step_kernel<G, RPL, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    lp_step_kernel<G, RPL, NC><<<grid, block, smem, stream>>>(st, state, S, days, slots, node_flow, col_flow, objective);
}

struct Variant {
    int G, RPL, NC;
    launch_fn launch;
    size_t (*smem_per_group)(int n_slots);
};

#define VARIANT(G, RPL, NC) \
    { G, RPL, NC, &launch_variant<G, RPL, NC>, &GroupSmem<G, RPL, NC>::bytes }

// ordered from the cheapest; the first one that fits (rows <= G*RPL, cols <= NC) is used
static const Variant kVariants[] = {
    VARIANT(8, 1, 4),   VARIANT(16, 1, 8),  VARIANT(32, 1, 8),  VARIANT(32, 1, 13),
    VARIANT(32, 1, 16), VARIANT(32, 1, 24), VARIANT(32, 2, 16), VARIANT(32, 2, 32),
};
static const int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

struct b200lp_handle {
    int device = 0;
    int S = 0;
    int variant = -1;
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

extern "C" {

const char *b200lp_last_error(void) { return g_last_error.c_str(); }
int b200lp_abi_version(void) { return B200LP_ABI_VERSION; }

int b200lp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int b200lp_create(const b200lp_structure *s, int32_t n_scenarios, int32_t device, b200lp_handle **out) {
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
    if (const char *force = getenv("B200LP_VARIANT")) {
        int v = atoi(force);
        if (v >= 0 && v < kNumVariants && m <= kVariants[v].G * kVariants[v].RPL && n <= kVariants[v].NC) variant = v;
    }
    if (variant < 0) {
        char buf[160];
        snprintf(buf, sizeof buf, "LP of %d rows x %d columns exceeds the register-resident simplex variants", m, n);
        return fail(B200LP_E_UNSUPPORTED, buf);
    }
    CUDA_TRY(cudaSetDevice(device));
    b200lp_handle *h = new b200lp_handle();
    h->device = device; h->S = n_scenarios; h->variant = variant;
    if (const char *b = getenv("B200LP_BLOCK")) {
        int v = atoi(b);
        if (v >= 32 && v <= 128 && v % 32 == 0) h->block = v;
    }
    const Variant &V = kVariants[variant];
    const int rows_cap = V.G * V.RPL, nc = V.NC;
    const size_t S = (size_t)n_scenarios;
    h->nnz = s->a_indptr[m];

    DevStructure &ds = h->ds;
    ds.m = m; ds.n = n; ds.N = N; ds.n_slots = s->n_slots; ds.n_consts = s->n_consts;
    ds.n_storage = s->n_storage; ds.n_dyn_a = s->n_dyn_a; ds.cost_clip = s->cost_clip;
    ds.rows_cap = rows_cap; ds.nc = nc;
    ds.cost_dynamic = 0;
    for (int k = 0; k < s->cost_indptr[n]; k++)
        if (s->cost_ref[k] >= 0) ds.cost_dynamic = 1;

    // validate refs
    auto ref_ok = [&](int ref) { return ref >= 0 ? ref < s->n_slots : (-ref - 1) < s->n_consts; };
    for (int i = 0; i < m; i++) {
        if (s->row_kind[i] == B200LP_ROW_FLOW) {
            if (!ref_ok(s->row_lb_ref[i]) || !ref_ok(s->row_ub_ref[i])) { delete h; return fail(B200LP_E_INVALID, "row ref out of range"); }
        } else if (s->row_lb_ref[i] < 0 || s->row_lb_ref[i] >= s->n_storage) { delete h; return fail(B200LP_E_INVALID, "storage index out of range"); }
        for (int k = s->a_indptr[i]; k < s->a_indptr[i + 1]; k++)
            if (s->a_col[k] < 0 || s->a_col[k] >= n) { delete h; return fail(B200LP_E_INVALID, "column index out of range"); }
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
    UP(row_kind, s->row_kind, m);
    UP(row_lb_ref, s->row_lb_ref, m);
    UP(row_ub_ref, s->row_ub_ref, m);
    UP(st_kind, s->st_kind, s->n_storage);
    UP(st_vol_ref, s->st_vol_ref, s->n_storage);
    UP(st_minv_ref, s->st_minv_ref, s->n_storage);
    UP(st_maxv_ref, s->st_maxv_ref, s->n_storage);
    UP(st_active_ref, s->st_active_ref, s->n_storage);
    UP(cost_indptr, s->cost_indptr, n + 1);
    UP(cost_ref, s->cost_ref, s->cost_indptr[n]);
    UP(cost_w, s->cost_w, s->cost_indptr[n]);
    UP(dyn_a_row, dyn_row.data(), s->n_dyn_a);
    UP(dyn_a_col, dyn_col.data(), s->n_dyn_a);
    UP(dyn_a_ref, s->dyn_a_ref, s->n_dyn_a);
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
    h->stats.n_rows = m; h->stats.n_cols = n; h->stats.n_nonzero = h->nnz; h->stats.kernel_variant = variant;
    *out = h;
    int rc = b200lp_reset(h, nullptr);
    if (rc != 0) { b200lp_destroy(h); *out = nullptr; return rc; }
    return B200LP_OK;
}

void b200lp_destroy(b200lp_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (void *p : h->owned) cudaFree(p);
    if (h->h_fail) cudaFreeHost(h->h_fail);
    if (h->h_counters) cudaFreeHost(h->h_counters);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int b200lp_reset(b200lp_handle *h, const double *initial_volume) {
    if (!h) return fail(B200LP_E_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
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
    if (h->ds.n_consts > 0) {
        CUDA_TRY(cudaMemcpyAsync(h->d_consts, consts, sizeof(double) * h->ds.n_consts, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    return B200LP_OK;
}

static int launch_step(b200lp_handle *h, double days, const double *d_slots, double *d_node_flow, double *d_col_flow,
                       double *d_objective) {
    const Variant &V = kVariants[h->variant];
    V.launch(h->ds, h->state, h->S, days, d_slots, d_node_flow, d_col_flow, d_objective, h->stream, h->block);
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
    return launch_step(h, days, d_slots ? d_slots : h->d_slots, d_node_flow, d_col_flow, d_objective);
}

int b200lp_step(b200lp_handle *h, double days, const double *slots, double *node_flow, double *col_flow,
                double *objective) {
    if (!h) return fail(B200LP_E_INVALID, "null handle");
    if (!(days > 0.0)) return fail(B200LP_E_INVALID, "days must be positive");
    if (h->ds.n_slots > 0 && !slots) return fail(B200LP_E_INVALID, "slots is NULL but the structure has input planes");
    CUDA_TRY(cudaSetDevice(h->device));
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
    h->stats.pivots = (int64_t)h->h_counters[0];
    h->stats.refactors = (int64_t)h->h_counters[1];
    *out = h->stats;
    return B200LP_OK;
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

}  // extern "C"
