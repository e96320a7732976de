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

#ifndef __CUDACC_RTC__ /* the run-time compiled kernel (pywr_b200/csrc/jit_run.inc) sees this text without a libc */
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* layout version of the structs below.  Program op codes are additive within a version: B200LP_OP_THRESHOLD .. B200LP_OP_STATE
   (13-16) were added to version 8 without touching a struct; a library that does not know an op rejects the program. */
#define B200LP_ABI_VERSION 8

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
    int32_t n_rows, n_cols, n_nonzero;
    int32_t kernel_variant;  /* index of the register variant; 100: CTA-per-scenario simplex; 200: PDHG; 300: revised simplex, inverse in HBM */
    int64_t pdhg_iterations; /* PDHG path: scenario-iterations since create                           */
    int32_t reduced_rows, reduced_nonzero; /* PDHG path: rows / non-zeros left by the structural presolve */
    int32_t run_kernel_jit;  /* 1: b200lp_run launches the structure-specialised kernel compiled at b200lp_load_program */
    int32_t reserved1;
} b200lp_stats;

const char *b200lp_last_error(void);
int b200lp_abi_version(void);
/* number of visible CUDA devices (0 when none: create will fail) */
int b200lp_device_count(void);

/* Build the engine for `n_scenarios` scenarios on CUDA device `device`. */
int b200lp_create(const b200lp_structure *s, int32_t n_scenarios, int32_t device, b200lp_handle **out);
/* Same with flags.  B200LP_CREATE_FOR_PROGRAM: the handle is meant for b200lp_load_program / b200lp_run; LPs beyond the
 * register-resident variants then take the large-LP path (which runs programs) instead of the CTA-per-scenario
 * shared-memory simplex (per-timestep b200lp_step only). */
#define B200LP_CREATE_FOR_PROGRAM 1u
int b200lp_create_ex(const b200lp_structure *s, int32_t n_scenarios, int32_t device, uint32_t flags, b200lp_handle **out);
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

/*
 * Engine selection is automatic (b200lp_create): register-resident simplex for LPs up to 64 rows x 32
 * columns, CTA-per-scenario shared-memory simplex up to 227 KB of dictionary, and beyond that the
 * large-LP path: structural presolve + scaling, then the revised simplex with every scenario's basis
 * inverse in HBM while those fit (kernel_variant 300; warm-started from the scenario's previous basis
 * like glp_simplex behind BasisManager, cython_glpk.pyx:197-232), else the batched restarted-PDHG
 * iterations (first-order, kernel_variant 200).  The environment variables B200LP_FORCE_CTA /
 * B200LP_FORCE_PDHG force the larger engines; B200LP_LARGE_ENGINE=rsx|pdhg picks the method of the
 * large-LP path.  b200lp_load_program / b200lp_run work on every engine except the CTA-per-scenario one.
 * Options of the PDHG path: "pdhg_tol" (relative KKT tolerance, default 1e-8: objectives within 1e-7 relative), "pdhg_max_iter"
 * (default 200000), "pdhg_check" (iterations between convergence tests, default 64), "pdhg_tail" (when no
 * more than this many scenarios still iterate after a batch of the streaming kernels they finish in the
 * CTA-per-scenario shared-memory kernel; default: no limit whenever one scenario's vectors fit 227 KB of
 * shared memory - measured 1.4x faster than streaming at 8,192 scenarios; 0: streaming kernels only).
 * The simplex engines accept and ignore pdhg_* options.
 */
int b200lp_set_option(b200lp_handle *h, const char *name, double value);

/* =====================================================================================
 * Device-resident run ("program"): the per-timestep inputs are produced ON the device, so that
 * a whole run needs no host round trip — parameter values feeding the LP (a3-a6 producers,
 * SURVEY.md section 8f rank 1), Storage.after mass balance (pywr/_core.pyx:1335-1361, a10) and the
 * built-in recorders' per-scenario accumulation (pywr/recorders/_recorders.pyx, a11-a13).
 *
 * Every timestep the ops are evaluated in order for every scenario; op k writes slot op_dst[k].
 * Integer arguments of op k are op_args[op_arg_ptr[k] .. op_arg_ptr[k+1]).
 * ===================================================================================== */
#define B200LP_OP_TABLE 1         /* args: table, offset.  values[clamp(t+offset)][column(scenario)]
                                     DataFrameParameter / ArrayIndexed(Scenario)Parameter / profiles
                                     pre-expanded over the run (pywr/parameters/_parameters.pyx:192-335,665-832) */
#define B200LP_OP_AGG 2           /* args: func, n, ref*n.  AggregatedParameter (:1439-1557)              */
#define B200LP_OP_CONTROL_CURVE 3 /* args: storage, n, curve_ref*n, value_ref*(n+1).
                                     ControlCurveParameter.value (_control_curves.pyx:501-524)           */
#define B200LP_OP_CC_INDEX 4      /* args: storage, n, curve_ref*n.  ControlCurveIndexParameter.index
                                     (_control_curves.pyx:345-364); the index is stored as a double      */
#define B200LP_OP_INDEXED 5       /* args: index_ref, n, ref*n.  IndexedArrayParameter.value (:1262-1312) */
#define B200LP_OP_NEG 6           /* args: ref.  NegativeParameter                                        */
#define B200LP_OP_MAX0 7          /* args: ref, threshold_ref.  MaxParameter: max(value, threshold)       */
#define B200LP_OP_MIN0 8          /* args: ref, threshold_ref.  MinParameter: min(value, threshold)       */
#define B200LP_OP_DIV 9           /* args: ref_a, ref_b.  a / b (AggregatedNode.get_factors_norm f0/fi,
                                     pywr/_core.pyx:921-939; DivisionParameter)                          */

#define B200LP_OP_STORAGE_RESET 10 /* args: storage, flag_ref, initial_volume_ref, initial_pc_ref.  Scheduled reset of a
                                     (virtual) storage in node.before(), i.e. BEFORE the parameters of the timestep are
                                     evaluated (AnnualVirtualStorage / SeasonalVirtualStorage / MonthlyVirtualStorage,
                                     pywr/nodes.py:596-757 -> Storage._reset_storage_only, pywr/_core.pyx:1294-1333):
                                     flag 1: volume = max_volume, pc = max_volume / max_volume;  flag 2: volume and pc =
                                     the initial ones;  0: nothing.  The op's slot receives the flag.                 */

#define B200LP_OP_CC_INTERP 11     /* args: storage, n, curve_ref*n, value_ref*(n+2).  ControlCurveInterpolatedParameter.value
                                     (_control_curves.pyx:176-216): linear interpolation between the values at 100 %, at each
                                     control curve and at 0 %; two coinciding levels return the lower level's value           */
#define B200LP_OP_CC_PIECEWISE 12  /* args: storage, n, curve_ref*n, upper_value_ref*(n+1), lower_value_ref*(n+1), minimum_ref,
                                     maximum_ref.  ControlCurvePiecewiseInterpolatedParameter.value (:269-292): within the band
                                     the storage is in, interpolation between that band's own pair of values (_interpolate, :10-24) */

#define B200LP_OP_THRESHOLD 13     /* args: source, x, predicate, threshold_ref, value0_ref, value1_ref.  AbstractThresholdParameter
                                     (pywr/parameters/_thresholds.pyx:19-131): value1 if predicate(X, threshold)
                                     else value0 (values 0 / 1: the parameter's index; a ratchet chains it with B200LP_OP_LAG on a
                                     recorder of its own slot).  source 0: X = the value behind ref x
                                     (ParameterThresholdParameter :294-318, calendar thresholds :362-395 through a table);
                                     source 1: X = the current volume of storage x (StorageThresholdParameter :134-157).
                                     predicate: B200LP_PRED_* */
#define B200LP_OP_INTERP 14        /* args: source, x, n (>= 2), xk_ref*n (ascending), yk_ref*n, below_ref, above_ref.  Piecewise-linear
                                     interpolation of X (source 0: the value behind ref x; source 1: the current volume of storage x):
                                     InterpolatedParameter / InterpolatedVolumeParameter (pywr/parameters/parameters.py:124-242,
                                     scipy interp1d kind="linear" = slope * (X - xk[j]) + yk[j]).  X below xk[0] / above xk[n-1] gives
                                     below / above (NaN refs reproduce bounds_error=True: the scenario then fails with B200LP_ST_NAN) */
#define B200LP_OP_LAG 15           /* args: recorder, initial_ref, steps (>= 1).  The sample array recorder `recorder` (any of the
                                     [T, S] kinds) took `steps` timesteps ago; initial_ref during the
                                     first `steps` timesteps.  A node's flow of the day before: FlowParameter
                                     (pywr/parameters/_parameters.pyx:1959-2008), AbstractNode._prev_flow as read by
                                     NodeThresholdParameter / MultipleThresholdIndexParameter (_thresholds.pyx:159-240); yesterday's
                                     deficit: DeficitParameter (:1910-1957); older flows: FlowDelayParameter (:2105-2163),
                                     RollingMeanFlowNodeParameter (:2191-2264).  The recorder's [T, S] array is the state: not
                                     available with B200LP_PROG_AGG_ONLY */
#define B200LP_OP_STATE 16         /* args: source, storage.  A storage's state as a value: source 1 its current volume, 2 its proportional
                                     volume as stored, 3 the same clamped to [0, 1] (Storage.get_current_pc, pywr/_core.pyx:1027-1039).
                                     StorageParameter (pywr/parameters/_parameters.pyx:2011-2044), Polynomial1DParameter on a storage */
#define B200LP_PRED_LT 0
#define B200LP_PRED_GT 1
#define B200LP_PRED_EQ 2
#define B200LP_PRED_LE 3
#define B200LP_PRED_GE 4

#define B200LP_AGG_SUM 0
#define B200LP_AGG_PRODUCT 1
#define B200LP_AGG_MIN 2
#define B200LP_AGG_MAX 3
#define B200LP_AGG_MEAN 4

/* table column selection */
#define B200LP_COL_NONE (-1)     /* single column                                                        */
#define B200LP_COL_SCENARIO (-2) /* one column per scenario combination (global_id)                      */
                                 /* >= 0: ScenarioIndex.indices[axis] (pywr/_core.pyx:108-128)           */

/* recorder kinds.  T = program steps.  "value" below is sum_k rec_w[k] * x[rec_col[k]], the node's
 * flow after commit (+ AggregatedNode / compound-node sums), restated as a linear form of the columns */
#define B200LP_REC_NODE_FLOW 1        /* [T][S]  value*factor                       (_recorders.pyx:611-617)   */
#define B200LP_REC_NODE_DEFICIT 2     /* [T][S]  max_flow - value                   (:680-688)                 */
#define B200LP_REC_NODE_SUPPLIED 3    /* [T][S]  value/max_flow, 1.0 on /0          (:723-734)                 */
#define B200LP_REC_NODE_CURTAIL 4     /* [T][S]  1 - value/max_flow, 0.0 on /0      (:769-779)                 */
#define B200LP_REC_STORAGE_VOLUME 5   /* [T][S]  volume after the mass balance      (:1175-1183)               */
#define B200LP_REC_STORAGE_PC 6       /* [T][S]  volume / max_volume                                           */
#define B200LP_REC_TOTAL_DEFICIT 7    /* [S]     += (max_flow - value)*days         (:1744-1753)               */
#define B200LP_REC_TOTAL_FLOW 8       /* [S]     += value*factor*days               (:1768-1775)               */
#define B200LP_REC_MEAN_FLOW 9        /* [S]     += value*factor   (the caller divides by T, :1795-1801)       */
#define B200LP_REC_DEFICIT_FREQ 10    /* [S]     += 1 if |value-max_flow| > 1e-6   (:1808-1822)               */
#define B200LP_REC_MIN_VOLUME 11      /* [S]     min(volume)                        (:1845-1852)               */

/* program flags.  AGG_ONLY: the array recorders keep only their running temporal aggregates (sum / min / max over the
 * timesteps, what Recorder.values() reduces _data to, pywr/recorders/_recorders.pyx:109-184,630-633); no [T][S] array is
 * allocated or written, b200lp_fetch_recorder on them fails with B200LP_E_UNSUPPORTED.  For optimisation batches and
 * ensembles that read objectives only (config 5: 11 GB of arrays otherwise). */
#define B200LP_PROG_AGG_ONLY 1

typedef struct b200lp_program {
    int32_t n_steps;     /* T: timesteps of the run                                                   */
    int32_t n_ops;
    int32_t n_tables;
    int32_t n_axes;      /* scenario axes (model.scenarios.scenarios)                                 */
    int32_t n_recorders;
    int32_t flags;       /* B200LP_PROG_*                                                              */
    const double *ts_days;      /* [T] Timestep.days (pywr/_core.pyx:276-301)                          */
    const int32_t *op_kind;     /* [n_ops]                                                             */
    const int32_t *op_dst;      /* [n_ops] slot written                                                */
    const int32_t *op_arg_ptr;  /* [n_ops+1]                                                           */
    const int32_t *op_args;
    const int32_t *table_rows;  /* [n_tables] 1 (time-invariant) or T                                  */
    const int32_t *table_cols;  /* [n_tables]                                                          */
    const int32_t *table_axis;  /* [n_tables] B200LP_COL_*                                             */
    const int64_t *table_offset; /* [n_tables] first element in table_data (row-major [rows][cols])    */
    const double *table_data;
    const int32_t *scen_index;  /* [n_axes][S] ScenarioIndex.indices                                   */
    /* storage mass balance.  Storage: the net inflow committed through StorageInput/StorageOutput
     * (_core.pyx:942-979) is exactly the activity of the storage's own LP row (+1 on columns ending in
     * its outputs, -1 on columns starting at its inputs, cython_glpk.pyx:578-601), which the engine
     * reads directly.  VirtualStorage (-sum factor*flow, :1483-1493): sum stflow_w * x[stflow_col].  */
    const int32_t *stflow_indptr; /* [n_storage+1] */
    const int32_t *stflow_col;
    const double *stflow_w;
    const double *initial_volume; /* [n_storage][S] Storage._volume after Model.reset                  */
    const double *initial_pc;     /* [n_storage][S] Storage._current_pc after Model.reset              */
    /* recorders */
    const int32_t *rec_kind;    /* [n_recorders]                                                       */
    const int32_t *rec_src;     /* [n_recorders] storage index for the storage kinds, else unused      */
    const int32_t *rec_ref;     /* [n_recorders] ref of the node's max_flow (deficit kinds)            */
    const double *rec_factor;   /* [n_recorders]                                                       */
    const int32_t *rec_row;     /* [n_recorders] >= 0: the node's flow is the activity of this LP row
                                   (non-storage row of a plain Input/Link/Output node: coefficient 1 on
                                   every column through the node, cython_glpk.pyx:515-550); -1: use the
                                   linear form below                                                   */
    const int32_t *rec_indptr;  /* [n_recorders+1] linear form of the node kinds                       */
    const int32_t *rec_col;
    const double *rec_w;
    /* RollingVirtualStorage (pywr/_core.pyx:1496-1536): st_roll_len[k] = timesteps - 1 of storage k (0: not rolling).
     * The engine keeps a ring buffer [st_roll_len[k]][S] per rolling storage, filled with st_roll_init[k] (the node's
     * _initial_utilisation) at reset; every timestep the entry of (t mod len) is added back to the volume
     * (Storage.after with adjustment: volume clamped to [min_volume, max_volume]) and then overwritten with the
     * volume used in this timestep (-flow * days).  Either pointer may be NULL when no storage is rolling. */
    const int32_t *st_roll_len;  /* [n_storage] */
    const double *st_roll_init;  /* [n_storage] */
} b200lp_program;

/* Load (copy) a program; allocates the recorder outputs on the device and resets the run state.
 * For the register-resident engine this also generates the CUDA source of a run kernel specialised on (structure,
 * program) - every index, row class, op and recorder kind an immediate, per-scenario scalars in registers - and compiles
 * it with NVRTC (libnvrtc is dlopen'ed; when it is missing, the model is not covered or B200LP_JIT=0 is set, the generic
 * run kernel is used: b200lp_stats.run_kernel_jit says which).  Same arithmetic, bit-identical results. */
int b200lp_load_program(b200lp_handle *h, const b200lp_program *p);
/* Diagnostic, needs no device: generate the specialised run kernel for (structure, program) and, with compile != 0,
 * compile it with NVRTC for sm_100a.  log (may be NULL) receives the compiler output including the ptxas resource
 * statistics.  B200LP_E_UNSUPPORTED: the generator does not cover this model (the generic run kernel would be used). */
int b200lp_jit_probe(const b200lp_structure *s, const b200lp_program *p, int32_t n_scenarios, int32_t compile, char *log, size_t cap);
/* Restart the loaded program: initial volumes, slack bases, zeroed recorders (Model.reset). */
int b200lp_program_reset(b200lp_handle *h);
/* Run timesteps [t0, t0+n) for all scenarios on the device.  Asynchronous on b200lp_stream();
 * errors of individual scenarios are reported by b200lp_program_status after a b200lp_sync. */
int b200lp_run(b200lp_handle *h, int32_t t0, int32_t n);
/*
 * One whole run with HOST buffers, pipelined: Model.reset, then the T timesteps in `n_chunks` time
 * slices; while slice c is being solved, the rows of the time-indexed tables that slice c+1 needs are
 * uploaded and the array-recorder rows slice c-1 produced are downloaded (three streams, both PCIe
 * directions busy at once).  table_src[k] (may be NULL, as may the array itself): new contents of table
 * k, same shape; rec_dst[r] (may be NULL): receives recorder r.  Pinned buffers (b200lp_host_alloc)
 * are needed for the copies to overlap.  Synchronous; results equal b200lp_update_table +
 * b200lp_program_reset + b200lp_run(0, T) + b200lp_fetch_recorder.
 */
int b200lp_run_pipelined(b200lp_handle *h, int32_t n_chunks, const double *const *table_src, double *const *rec_dst);
/* Number of failed scenarios so far, the first failing scenario (or -1), its status and timestep. */
int b200lp_program_status(b200lp_handle *h, int32_t *n_failed, int32_t *scenario, int32_t *code, int32_t *timestep);
/* Replace the contents of table `table` (same shape) from host memory — pinned memory is copied
 * at full PCIe rate, asynchronously on b200lp_stream().  New inflow scenarios for the next run
 * without rebuilding the program (the optimisation / ensemble outer loop). */
int b200lp_update_table(b200lp_handle *h, int32_t table, const double *data);
/* Copy recorder `rec` to the host: [T][S] doubles for the array kinds, [S] for the others. */
int b200lp_fetch_recorder(b200lp_handle *h, int32_t rec, double *out);
/* The same into a column block of a wider host array: row t of the recorder goes to out[t * ld .. t * ld + S), ld >= S
 * (one cudaMemcpy2DAsync on b200lp_stream(); NOT synchronised: call b200lp_sync before reading).  This is how a host that
 * shards one scenario batch over several handles / GPUs assembles the reference's _data[nts, ncomb] layout
 * (pywr/recorders/_recorders.pyx:574-617) without a staging copy: `out` points at the shard's first column. */
int b200lp_fetch_recorder_strided(b200lp_handle *h, int32_t rec, double *out, int64_t ld);
/* Temporal aggregates of array recorder `rec`, accumulated on the device in timestep order while the run writes the
 * array: sum, min and max over the timesteps executed since the last reset, [S] doubles each (any pointer may be NULL).
 * The caller derives Recorder.values() for the recorder's temporal agg_func: "sum", "min", "max", "mean" = sum / T
 * (Aggregator.aggregate_2d(axis=0), pywr/recorders/_recorders.pyx:109-184,630-633) without moving the [T][S] array. */
int b200lp_fetch_recorder_agg(b200lp_handle *h, int32_t rec, double *sum, double *min, double *max);
/* Device pointer of recorder `rec` (for NCCL gathers by the caller) and its element count. */
void *b200lp_recorder_device_ptr(b200lp_handle *h, int32_t rec, int64_t *count);
/* Column values x [n][S] of the last executed timestep (host mirrors of node flows) and the current
 * volumes / proportions [n_storage][S]; any pointer may be NULL. */
int b200lp_fetch_state(b200lp_handle *h, double *col_flow, double *volume, double *pc);

/* CUDA-event stopwatch on the engine's stream: mark(0) = start, mark(1) = stop; elapsed_ms
 * synchronises on the stop event.  (torch.cuda.Event only sees torch's own streams.) */
int b200lp_mark(b200lp_handle *h, int32_t which);
int b200lp_elapsed_ms(b200lp_handle *h, double *ms);

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
