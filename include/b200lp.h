/*
 * b200lp.h — C-ABI of the B200 scenario-batched LP engine for Pywr.
 *
 * This is the drop-in boundary for the one hot path of pywr/pywr that this repository
 * accelerates: the per-timestep, per-scenario loop of CythonGLPKSolver.solve
 * (reference: pywr/solvers/cython_glpk.pyx:821-1095 for the route formulation and
 * :1649-1918 for the edge formulation), plus flow commit (pywr/_core.pyx:481-489,942-1004),
 * storage mass balance (pywr/_core.pyx:1335-1361) and the built-in array / deficit / storage
 * recorders (pywr/recorders/_recorders.pyx:574-780,1151-1184,1740-1853).
 *
 * Conventions
 *   - plain C, no torch / Python types; every pointer argument is caller-owned HOST memory unless
 *     its name starts with d_ (device pointer);
 *   - every function returns 0 on success and a negative B200LP_E_* code on failure;
 *     b200lp_last_error() returns a human-readable message for the calling thread;
 *   - all floating point is IEEE double, all indices are int32;
 *   - per-scenario arrays are laid out [item][scenario] with the scenario index contiguous
 *     ("scenario-major" planes, one plane per item), scenario = ScenarioIndex.global_id
 *     (pywr/_core.pyx:108-128,210-240);
 *   - there is NO CPU fallback: b200lp_create fails with B200LP_E_NO_DEVICE without a CUDA device.
 *
 * Value references ("refs"): wherever the LP needs a per-scenario number (a row bound, a node
 * cost, a storage capacity, an aggregated-node factor) the structure holds an int32 ref:
 *     ref >= 0  -> slot `ref` of the per-step input planes  slots[ref][scenario]
 *     ref <  0  -> constant  consts[-ref-1]
 * A "slot" therefore plays the role of one Node.get_min_flow / get_max_flow / get_cost /
 * Storage.get_max_volume ... call of the reference, evaluated for all scenarios at once
 * (pywr/_core.pyx:566-574,618-626,681-689,1202-1238: get_all_*).
 */
#ifndef B200LP_H
#define B200LP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200LP_ABI_VERSION 3

/* error codes */
#define B200LP_OK 0
#define B200LP_E_INVALID (-1)     /* bad argument / inconsistent structure            */
#define B200LP_E_NO_DEVICE (-2)   /* no CUDA device: there is no CPU fallback         */
#define B200LP_E_CUDA (-3)        /* a CUDA runtime call failed                       */
#define B200LP_E_UNSUPPORTED (-4) /* LP too large for every compiled kernel variant   */
#define B200LP_E_SOLVE (-5)       /* at least one scenario did not reach an optimum   */

/* per-scenario solve status (b200lp_status) */
#define B200LP_ST_OPTIMAL 0
#define B200LP_ST_INFEASIBLE 1 /* "no primal feasible solution"  -> RuntimeError, cython_glpk.pyx:1060 */
#define B200LP_ST_UNBOUNDED 2  /* "no dual feasible solution"                                          */
#define B200LP_ST_ITERLIMIT 3  /* iteration limit exceeded                                             */
#define B200LP_ST_NAN 4        /* NaN in a bound / cost / coefficient -> GLPKError, :286-335           */
#define B200LP_ST_SINGULAR 5   /* basis could not be re-factorised even from the slack basis           */

/* row kinds: how the bounds of LP row i are produced every timestep */
#define B200LP_ROW_FLOW 0     /* lb = ref(row_lb_ref), ub = ref(row_ub_ref)   (:932-959, a5)          */
#define B200LP_ROW_STORAGE 1  /* row_lb_ref = storage index k; bounds from volume (:965-988, a6)      */

/* storage kinds */
#define B200LP_ST_KIND_STORAGE 0
#define B200LP_ST_KIND_VIRTUAL 1 /* has an `active` flag; inactive -> free row (:1000-1004)           */

/* ref meaning "the engine's own device-resident storage volume" (st_vol_ref only) */
#define B200LP_REF_STATE INT32_MIN

/*
 * The LP structure shared by every scenario.  Built once at Solver.setup
 * (replaces cython_glpk.pyx:388-755 / :1148-1567).  All arrays are copied by b200lp_create.
 */
typedef struct b200lp_structure {
    int32_t abi_version;   /* = B200LP_ABI_VERSION                                                   */
    int32_t n_rows;        /* m                                                                       */
    int32_t n_cols;        /* n: routes (route form) or edges (edge form); all columns are x >= 0     */
    int32_t n_nodes;       /* N: graph nodes sorted by fully_qualified_name (:409)                    */
    int32_t n_slots;       /* planes in the per-step input                                            */
    int32_t n_consts;
    int32_t n_storage;     /* Storage + VirtualStorage rows                                           */
    int32_t n_dyn_a;       /* matrix non-zeros rewritten per scenario-step (:903-926, a4)             */
    int32_t cost_clip;     /* 1: |c_j| < 1e-8 -> 0 (route form :892); 0: edge form                    */
    int32_t reserved0;

    /* constraint matrix, CSR: rows in the reference's order (non-storage, cross-domain,
     * [link mass balance], storage, virtual storage, aggregated factors, aggregated min/max) */
    const int32_t *a_indptr; /* [m+1] */
    const int32_t *a_col;    /* [nnz] */
    const double *a_val;     /* [nnz] */

    const int32_t *row_kind;   /* [m] B200LP_ROW_*                                                    */
    const int32_t *row_lb_ref; /* [m] ref, or storage index for B200LP_ROW_STORAGE                    */
    const int32_t *row_ub_ref; /* [m] ref (unused for storage rows)                                   */

    const int32_t *st_kind;       /* [n_storage]                                                      */
    const int32_t *st_vol_ref;    /* [n_storage] slot carrying Storage._volume, or B200LP_REF_STATE   */
    const int32_t *st_minv_ref;   /* [n_storage] ref: get_min_volume                                  */
    const int32_t *st_maxv_ref;   /* [n_storage] ref: get_max_volume                                  */
    const int32_t *st_active_ref; /* [n_storage] ref: != 0 -> active (virtual storage only)           */

    /* objective: c_j = sum_k cost_w[k] * ref(cost_ref[k]),  k in [cost_indptr[j], cost_indptr[j+1]) */
    const int32_t *cost_indptr; /* [n+1] */
    const int32_t *cost_ref;
    const double *cost_w;

    /* dynamic matrix entries: a_val[dyn_a_nz[k]] = dyn_a_scale[k] * ref(dyn_a_ref[k])               */
    const int32_t *dyn_a_nz;
    const int32_t *dyn_a_ref;
    const double *dyn_a_scale;

    /* result extraction: node_flow[v] = sum_k flow_w[k] * x[flow_col[k]] (:1078-1088 / :1899-1911)  */
    const int32_t *flow_indptr; /* [N+1] */
    const int32_t *flow_col;
    const double *flow_w;

    const double *consts; /* [n_consts] */
} b200lp_structure;

typedef struct b200lp_handle b200lp_handle;

/* counters filled by b200lp_get_stats (role of CythonGLPKSolver.stats, :742-755) */
typedef struct b200lp_stats {
    int64_t steps;           /* timesteps solved since create                                        */
    int64_t scenario_steps;  /* steps * scenarios                                                    */
    int64_t pivots;          /* simplex pivots over all scenarios                                    */
    int64_t refactors;       /* basis re-factorisations over all scenarios                           */
    int64_t kernel_launches; /* launches of this library's kernels                                   */
    double h2d_bytes, d2h_bytes;
    double device_ms;        /* CUDA-event time of the solve kernels                                 */
    int32_t n_rows, n_cols, n_nonzero, kernel_variant;
} b200lp_stats;

const char *b200lp_last_error(void);
int b200lp_abi_version(void);
/* number of visible CUDA devices (0 when none: create will fail) */
int b200lp_device_count(void);

/* Build the engine for `n_scenarios` scenarios on CUDA device `device`. */
int b200lp_create(const b200lp_structure *s, int32_t n_scenarios, int32_t device, b200lp_handle **out);
void b200lp_destroy(b200lp_handle *h);

/*
 * Solver.reset(): forget every scenario's basis (the next solve crashes a new one, role of
 * is_first_solve / glp_adv_basis, :218-224,757-759) and load the device-resident storage
 * volumes ([n_storage][S], may be NULL when every st_vol_ref is a slot).
 */
int b200lp_reset(b200lp_handle *h, const double *initial_volume);

/*
 * One timestep for all scenarios (the loop of CythonGLPKSolver.solve, :821-831):
 * upload slots [n_slots][S], update bounds / costs / dynamic coefficients, warm-started solve,
 * extract flows.  Outputs (any may be NULL): node_flow [N][S] (what the reference passes to
 * node.commit, :1091-1093), col_flow [n][S] (route_flows_arr, :1069-1075), objective [S].
 * Synchronous: on return the outputs are valid.  Returns B200LP_E_SOLVE if any scenario failed;
 * details through b200lp_status.
 */
int b200lp_step(b200lp_handle *h, double days, const double *slots, double *node_flow, double *col_flow,
                double *objective);

/* Same with device-resident planes (inputs already in HBM); nothing is copied, nothing synchronised.
 * Work is queued on the engine's stream (b200lp_stream). */
int b200lp_step_device(b200lp_handle *h, double days, const double *d_slots, double *d_node_flow,
                       double *d_col_flow, double *d_objective);

/* Replace the constant pool (consts[n_consts]): used when a literal node attribute (a max_flow set
 * to a new number between runs or steps) changes without a structural change. */
int b200lp_set_consts(b200lp_handle *h, const double *consts);

/* first failing scenario of the last step (or -1) and its B200LP_ST_* code */
int b200lp_status(b200lp_handle *h, int32_t *first_bad_scenario, int32_t *code);

/* basis of every scenario as GLPK-style status codes (1 basic, 2 at lower, 3 at upper, 4 free,
 * 5 fixed), row_stat [S][m], col_stat [S][n] — the layout of BasisManager (:197-216) */
int b200lp_get_basis(b200lp_handle *h, int32_t *row_stat, int32_t *col_stat);

/* device-resident storage volumes [n_storage][S] */
int b200lp_get_volume(b200lp_handle *h, double *volume);
int b200lp_set_volume(b200lp_handle *h, const double *volume);

int b200lp_get_stats(b200lp_handle *h, b200lp_stats *out);

/* the cudaStream_t all work of this handle is queued on (for CUDA-event timing by the caller) */
void *b200lp_stream(b200lp_handle *h);
int b200lp_sync(b200lp_handle *h);

/* pinned host memory for the gather buffers (cudaHostAlloc / cudaFreeHost) and plain device memory */
void *b200lp_host_alloc(size_t bytes);
void b200lp_host_free(void *p);
void *b200lp_device_alloc(size_t bytes);
void b200lp_device_free(void *p);
int b200lp_memcpy_h2d(b200lp_handle *h, void *d_dst, const void *src, size_t bytes);
int b200lp_memcpy_d2h(b200lp_handle *h, void *dst, const void *d_src, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* B200LP_H */
