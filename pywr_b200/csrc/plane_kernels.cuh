// plane_kernels.cuh / plane_program.inc — device-resident run (b200lp_load_program / b200lp_run) for the large-LP path (revised simplex
// or PDHG): the same program the run kernel of the register engines executes inside one launch
// (lp_kernel.cuh::lp_run_kernel), here as three small kernels around the batched solve of every timestep, one thread
// per scenario over planes [item][S]:
//   plane_ops_kernel    Model.before: table look-ups and parameter ops into the slot planes
//                       (pywr/_model.pyx:825-849; ops B200LP_OP_*, same interpreter as the run kernel: eval_ops)
//   (assembly + solve)  pdhg_step_device: bounds / costs from the planes, revised simplex or PDHG, un-scaling
//   plane_after_kernel  Storage.after mass balance (pywr/_core.pyx:1335-1361, virtual :1483-1493, rolling :1522-1536)
//                       and the recorders' after() (pywr/recorders/_recorders.pyx:611-617,680-688,723-734,769-779,
//                       1175-1183,1744-1822,1845-1852)
// The value vector V = [slots | constants | volumes | pc | built-in constants] of the run kernel becomes a stack of
// planes; the volume and pc planes ARE the engine's storage state (h->state.vol, dp.pc alias them).  A node's flow is
// the activity of its LP row, recomputed from the column values (the run kernel reads it from the simplex).

#pragma once

struct PlaneProg {
    int S, ld, n_slots, n_consts, n_storage, vol_base, pc_base, k_base;
    double *V;  // [n_val + K_COUNT][S]
    const int *a_indptr, *a_col;  // ORIGINAL rows
    const double *a_val;
    const int *st_kind, *st_row, *st_maxv_idx, *st_minv_idx, *st_active_idx;
    const double *xb;  // [n][ld] column values of the last solve
    int *status;       // [ld] status of the last solve
    int *dead;         // [S] 0, or the B200LP_ST_* code the scenario failed with at an earlier timestep
    int *fail;         // [2] number of failed scenarios, min((scenario << 4) | code)
};

struct PlaneView {
    double *base;
    size_t stride;
    __device__ __forceinline__ double &operator[](const int i) const { return base[(size_t)i * stride]; }
};

__global__ void plane_fill_consts(const PlaneProg Q, const double *__restrict__ consts) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= Q.S) return;
    const size_t S = (size_t)Q.S;
    for (int k = 0; k < Q.n_consts; k++) Q.V[(size_t)(Q.n_slots + k) * S + s] = __ldg(&consts[k]);
    double *K = Q.V + (size_t)Q.k_base * S + s;
    K[K_SCRATCH * S] = 0.0; K[K_ZERO * S] = 0.0; K[K_ONE * S] = 1.0; K[K_TWO * S] = 2.0; K[K_PINF * S] = INFINITY; K[K_NINF * S] = -INFINITY;
}

__global__ void plane_ops_kernel(const PlaneProg Q, const DevProgram pg, const int t) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= Q.S || Q.dead[s] != 0) return;
    const PlaneView V{Q.V + s, (size_t)Q.S};
    for (int j = 0; j < pg.n_table_ops; j++) {
        const int tb = __ldg(&pg.tab_table[j]);
        const int ax = __ldg(&pg.table_axis[tb]);
        const int col = ax == -1 ? 0 : (ax == -2 ? s : __ldg(&pg.scen_index[(size_t)ax * Q.S + s]));
        TableDesc q;
        q.ptr = pg.table_data + __ldg(&pg.table_offset[tb]) + col;
        q.last = __ldg(&pg.table_rows[tb]) - 1; q.stride = __ldg(&pg.table_cols[tb]); q.off = __ldg(&pg.tab_offset[j]);
        V[__ldg(&pg.tab_dst[j])] = q.at(t);
    }
    if (pg.n_code > 0) eval_ops(pg.code, pg.n_code, V, Q.pc_base, LagCtx{pg.rec_out, t, Q.S, s});
}

__global__ void plane_clear_status(const PlaneProg Q) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < Q.ld) Q.status[s] = s < Q.S ? Q.dead[s] : 0;
}

__global__ void plane_after_kernel(const PlaneProg Q, const DevProgram pg, const int t, const double days) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= Q.S || Q.dead[s] != 0) return;
    const int st = Q.status[s];
    if (st != B200LP_ST_OPTIMAL) {  // the scenario stops here, like a failed scenario of the run kernel
        Q.dead[s] = st;
        pg.fail_step[s] = t;
        atomicAdd(&Q.fail[0], 1);
        atomicMin(&Q.fail[1], (s << 4) | st);
        return;
    }
    const size_t S = (size_t)Q.S;
    const PlaneView V{Q.V + s, S};
    const double *x = Q.xb + s;
    const size_t ld = (size_t)Q.ld;
    auto row_activity = [&](const int i) {
        double acc = 0.0;
        for (int e = __ldg(&Q.a_indptr[i]); e < __ldg(&Q.a_indptr[i + 1]); e++) acc = fma(__ldg(&Q.a_val[e]), x[(size_t)__ldg(&Q.a_col[e]) * ld], acc);
        return acc;
    };
    // ---- Storage.after ------------------------------------------------------------------------------------------
    for (int k = 0; k < Q.n_storage; k++) {
        const int kind = __ldg(&Q.st_kind[k]);
        if (kind == B200LP_ST_KIND_VIRTUAL && !(V[__ldg(&Q.st_active_idx[k])] != 0.0)) continue;
        double flow = 0.0;
        if (kind == B200LP_ST_KIND_STORAGE) flow = row_activity(__ldg(&Q.st_row[k]));
        else
            for (int e = __ldg(&pg.stflow_indptr[k]); e < __ldg(&pg.stflow_indptr[k + 1]); e++)
                flow = fma(__ldg(&pg.stflow_w[e]), x[(size_t)__ldg(&pg.stflow_col[e]) * ld], flow);
        double v = V[Q.vol_base + k] + flow * days;
        const double mxv = V[__ldg(&Q.st_maxv_idx[k])], mnv = V[__ldg(&Q.st_minv_idx[k])];
        if (pg.roll_len != nullptr && __ldg(&pg.roll_len[k]) > 0) {
            double *mem = pg.roll_mem + ((size_t)__ldg(&pg.roll_off[k]) + (size_t)(t % __ldg(&pg.roll_len[k]))) * S + s;
            v += *mem;
            v = fmin(mxv, fmax(mnv, v));
            *mem = -flow * days;
        }
        if (fabs(v - mxv) < 1e-6) v = mxv;
        if (fabs(v - mnv) < 1e-6) v = mnv;
        V[Q.vol_base + k] = v;
        V[Q.pc_base + k] = (mxv == 0.0) ? NAN : v / mxv;
    }
    // ---- recorders ----------------------------------------------------------------------------------------------
    for (int r = 0; r < pg.n_recorders; r++) {
        const int kind = __ldg(&pg.rec_kind[r]), row = __ldg(&pg.rec_row[r]), idx = __ldg(&pg.rec_idx[r]);
        const int src = __ldg(&pg.rec_src[r]);
        const double f = __ldg(&pg.rec_factor[r]);
        double *out = pg.rec_out[r];
        double value = 0.0;
        const bool node_kind = !(kind == B200LP_REC_STORAGE_VOLUME || kind == B200LP_REC_STORAGE_PC || kind == B200LP_REC_MIN_VOLUME);
        if (node_kind) {
            if (row >= 0) value = row_activity(row);
            else
                for (int e = __ldg(&pg.rec_indptr[r]); e < __ldg(&pg.rec_indptr[r + 1]); e++)
                    value = fma(__ldg(&pg.rec_w[e]), x[(size_t)__ldg(&pg.rec_col[e]) * ld], value);
        }
        const size_t ts_at = (size_t)t * S + s;
        if (kind >= B200LP_REC_NODE_FLOW && kind <= B200LP_REC_STORAGE_PC) {
            double o = 0.0;
            switch (kind) {
            case B200LP_REC_NODE_FLOW: o = value * f; break;
            case B200LP_REC_NODE_DEFICIT: o = V[idx] - value; break;
            case B200LP_REC_NODE_SUPPLIED: { const double mf = V[idx]; o = mf == 0.0 ? 1.0 : value / mf; } break;
            case B200LP_REC_NODE_CURTAIL: { const double mf = V[idx]; o = mf == 0.0 ? 0.0 : 1.0 - value / mf; } break;
            case B200LP_REC_STORAGE_VOLUME: o = V[Q.vol_base + src]; break;
            default: o = V[Q.pc_base + src]; break;
            }
            if (out) out[ts_at] = o;
            double *ag = pg.rec_agg + (size_t)r * 3 * S + s;   // running sum / min / max over the timesteps (values())
            agg_sample(ag[0], ag[S], ag[2 * S], o);
            continue;
        }
        switch (kind) {
        case B200LP_REC_TOTAL_DEFICIT: out[s] += (V[idx] - value) * days; break;
        case B200LP_REC_TOTAL_FLOW: out[s] += value * f * days; break;
        case B200LP_REC_MEAN_FLOW: out[s] += value * f; break;
        case B200LP_REC_DEFICIT_FREQ: if (fabs(value - V[idx]) > 1e-6) out[s] += 1.0; break;
        case B200LP_REC_MIN_VOLUME: { const double v = V[Q.vol_base + src]; if (v < out[s]) out[s] = v; } break;
        default: break;
        }
    }
}

// column values of the last executed timestep, [n][S] (b200lp_fetch_state)
__global__ void plane_last_x(const PlaneProg Q, const int n, double *__restrict__ last_x) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (s >= Q.S || j >= n || Q.dead[s] != 0) return;
    last_x[(size_t)j * Q.S + s] = Q.xb[(size_t)j * Q.ld + s];
}

